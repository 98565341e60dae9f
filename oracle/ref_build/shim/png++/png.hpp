// Stub of the png++ API surface that src/texture.cpp:125-139,189-208 touches.
// The oracle never reads or writes PNG files: construction from a path yields an
// empty image, write() is a no-op.  Test infrastructure only.
#pragma once
#include <string>
#include <vector>
#include <cstdint>
#include <fstream>
#include <iostream>
namespace png {
typedef uint32_t uint_32;
struct rgb_pixel {
    unsigned char red, green, blue;
    rgb_pixel() : red(0), green(0), blue(0) {}
    template <class A, class B, class C> rgb_pixel(A r, B g, C b) : red((unsigned char)r), green((unsigned char)g), blue((unsigned char)b) {}
};
template <class P> class image {
    uint_32 w_, h_;
    std::vector<std::vector<P>> rows_;
public:
    image(uint_32 w, uint_32 h) : w_(w), h_(h), rows_(h, std::vector<P>(w)) {}
    explicit image(const std::string&) : w_(0), h_(0) {}
    uint_32 get_width() const { return w_; }
    uint_32 get_height() const { return h_; }
    P get_pixel(uint_32 x, uint_32 y) const { return rows_[y][x]; }
    std::vector<P>& operator[](uint_32 y) { return rows_[y]; }
    void write(const std::string&) const {}
};
}
