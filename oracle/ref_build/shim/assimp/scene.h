// Stand-in for the assimp data structures that src/scene.cpp:58-207 and
// src/bxdf/bxdf.cpp:88-184 read.  Not assimp: plain structs that the oracle
// harness fills from a scene pack so the reference's own LoadAiSceneMeshes /
// LoadFromAiMaterial code paths ingest it.  There is no importer (ReadFile fails).
// Test infrastructure only.
#pragma once
#include <string>
#include <cstring>

struct aiVector3D { float x, y, z; };
struct aiColor3D { float r, g, b; aiColor3D() : r(0), g(0), b(0) {} };
struct aiString {
    std::string s;
    const char* C_Str() const { return s.c_str(); }
};
struct aiMatrix4x4 {
    float m[4][4];
    aiMatrix4x4() { std::memset(m, 0, sizeof m); m[0][0] = m[1][1] = m[2][2] = m[3][3] = 1.0f; }
    float* operator[](unsigned i) { return m[i]; }
    const float* operator[](unsigned i) const { return m[i]; }
};
struct aiFace { unsigned int mNumIndices; unsigned int* mIndices; };
struct aiMesh {
    unsigned int mNumVertices = 0, mNumFaces = 0, mMaterialIndex = 0;
    aiVector3D* mVertices = nullptr;
    aiVector3D* mNormals = nullptr;
    aiVector3D* mTangents = nullptr;
    aiVector3D* mTextureCoords[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    aiFace* mFaces = nullptr;
};
struct aiNode {
    aiMatrix4x4 mTransformation;
    unsigned int mNumMeshes = 0; unsigned int* mMeshes = nullptr;
    unsigned int mNumChildren = 0; aiNode** mChildren = nullptr;
};
enum aiTextureType { aiTextureType_DIFFUSE = 1, aiTextureType_SPECULAR = 2, aiTextureType_HEIGHT = 5 };
enum aiReturn { aiReturn_SUCCESS = 0, aiReturn_FAILURE = -1 };

#define AI_MATKEY_NAME "?mat.name", 0, 0
#define AI_MATKEY_COLOR_DIFFUSE "$clr.diffuse", 0, 0
#define AI_MATKEY_COLOR_SPECULAR "$clr.specular", 0, 0
#define AI_MATKEY_COLOR_EMISSIVE "$clr.emissive", 0, 0
#define AI_MATKEY_SHININESS "$mat.shininess", 0, 0

struct aiMaterial {
    std::string name;
    aiColor3D diffuse, specular, emissive;
    float shininess = 0.0f;
    std::string tex_diffuse, tex_specular, tex_height;
    aiReturn Get(const char* key, unsigned, unsigned, aiString& out) const {
        if (!std::strcmp(key, "?mat.name")) { out.s = name; return aiReturn_SUCCESS; }
        return aiReturn_FAILURE;
    }
    aiReturn Get(const char* key, unsigned, unsigned, aiColor3D& out) const {
        if (!std::strcmp(key, "$clr.diffuse")) { out = diffuse; return aiReturn_SUCCESS; }
        if (!std::strcmp(key, "$clr.specular")) { out = specular; return aiReturn_SUCCESS; }
        if (!std::strcmp(key, "$clr.emissive")) { out = emissive; return aiReturn_SUCCESS; }
        return aiReturn_FAILURE;
    }
    aiReturn Get(const char* key, unsigned, unsigned, float& out) const {
        if (!std::strcmp(key, "$mat.shininess")) { out = shininess; return aiReturn_SUCCESS; }
        return aiReturn_FAILURE;
    }
    const std::string& tex(aiTextureType t) const {
        return t == aiTextureType_DIFFUSE ? tex_diffuse : (t == aiTextureType_SPECULAR ? tex_specular : tex_height);
    }
    unsigned int GetTextureCount(aiTextureType t) const { return tex(t).empty() ? 0u : 1u; }
    aiReturn GetTexture(aiTextureType t, unsigned, aiString* path) const { path->s = tex(t); return aiReturn_SUCCESS; }
};
struct aiScene {
    unsigned int mNumMeshes = 0; aiMesh** mMeshes = nullptr;
    unsigned int mNumMaterials = 0; aiMaterial** mMaterials = nullptr;
    aiNode* mRootNode = nullptr;
};
