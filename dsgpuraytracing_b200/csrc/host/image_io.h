// image_io.h -- environment-map readers of the host side: OpenEXR (what the reference's -e option loads through its
// vendored tinyexr, src/main.cpp:30-67) and the portable float map (.pfm).
#pragma once
#include <string>

#include "pathtracer.h"

namespace dsrt_host {

// Scan-line OpenEXR, single part, channels R, G, B (HALF, FLOAT or UINT; other channels are ignored), compression
// NONE, RLE, ZIPS, ZIP or PIZ.  Rows are returned top first, which is the order EnvironmentLight indexes them
// (environment_light.cpp:24-27: row 0 = theta 0 = +y pole).  Kept quirk of the reference's tinyexr: a file whose lineOrder
// is DECREASING_Y comes out mirrored vertically.
bool load_exr(const std::string& path, HDRImageBuffer& img, std::string& err);
// "PF\n<w> <h>\n<-scale>\n" + h rows of w RGB float triplets, bottom row first, little endian
bool load_pfm(const std::string& path, HDRImageBuffer& img, std::string& err);
// picks the reader from the file's magic number
bool load_envmap(const std::string& path, HDRImageBuffer& img, std::string& err);

}  // namespace dsrt_host
