// Stub of the OpenEXR API surface used by src/texture.cpp:356-374.  The oracle
// build does not write EXR files (writePixels is a no-op); images are read back
// through EXRTexture::GetPixel instead.  Test infrastructure only.
#pragma once
namespace Imf {
struct Rgba { float r, g, b, a; };
enum RgbaChannels { WRITE_RGBA = 0x0f };
class RgbaOutputFile {
public:
    RgbaOutputFile(const char*, int, int, RgbaChannels) {}
    void setFrameBuffer(const Rgba*, int, int) {}
    void writePixels(int) {}
};
}
