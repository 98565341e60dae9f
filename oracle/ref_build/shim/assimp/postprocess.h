#pragma once
enum aiPostProcessSteps {
    aiProcess_CalcTangentSpace = 0x1, aiProcess_JoinIdenticalVertices = 0x2, aiProcess_Triangulate = 0x8,
    aiProcess_GenNormals = 0x20, aiProcess_GenSmoothNormals = 0x40, aiProcess_FindDegenerates = 0x10000,
    aiProcess_GenUVCoords = 0x40000
};
