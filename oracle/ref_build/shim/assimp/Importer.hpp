// Importer stand-in for src/config.cpp:196-228: there is no mesh importer in the
// oracle build, ReadFile always fails (the caller throws ConfigFileException).
#pragma once
#include <string>
#include "scene.h"
#define AI_CONFIG_PP_SBP_REMOVE "PP_SBP_REMOVE"
enum aiPrimitiveType { aiPrimitiveType_POINT = 1, aiPrimitiveType_LINE = 2 };
namespace Assimp {
class Importer {
public:
    bool SetPropertyInteger(const char*, int, void*) { return true; }
    const aiScene* ReadFile(const std::string&, unsigned int) { return nullptr; }
    const aiScene* ApplyPostProcessing(unsigned int) { return nullptr; }
    const char* GetErrorString() const { return "mesh import is not available in the oracle build"; }
};
}
