/* Stub of the libjpeg API surface used by src/texture.cpp:210-291.  Decoding is
 * not available in the oracle build: jpeg_read_header reports a 0x0 image with 0
 * components, so CreateNewFromJPEG returns nullptr.  Test infrastructure only. */
#pragma once
#include <cstdio>
#ifndef TRUE
#define TRUE 1
#endif
struct jpeg_error_mgr { int dummy; };
struct jpeg_decompress_struct {
    jpeg_error_mgr* err;
    unsigned int output_width, output_height, output_scanline;
    int num_components;
};
inline jpeg_error_mgr* jpeg_std_error(jpeg_error_mgr* e) { return e; }
inline void jpeg_create_decompress(jpeg_decompress_struct* i) { i->output_width = i->output_height = i->output_scanline = 0; i->num_components = 0; }
inline void jpeg_stdio_src(jpeg_decompress_struct*, FILE*) {}
inline int jpeg_read_header(jpeg_decompress_struct*, int) { return 0; }
inline int jpeg_start_decompress(jpeg_decompress_struct*) { return 0; }
inline unsigned int jpeg_read_scanlines(jpeg_decompress_struct* i, unsigned char**, unsigned int n) { i->output_scanline += n; return n; }
inline int jpeg_finish_decompress(jpeg_decompress_struct*) { return 0; }
