// image_io.h -- 8-bit image files as the reference's device keeps them after loading (the part of IG::Image this layer needs).
//
// The generated stage text does not hold pixels, it names FILES: `device.load_packed_image_by_id(<resource id>, <channels>, <linear>)`
// (src/runtime/pattern/ImagePattern.cpp:57-66; the id indexes IRenderDevice::SceneSettings::resource_map). The reference's device decodes the
// file with stb_image through IG::Image::loadAsPacked (src/runtime/Image.cpp:714-810; src/device/Device.cpp:735-799): rows flipped bottom-up,
// grey files kept as one byte per pixel, everything else as RGBA bytes, and -- unless the texture says `linear` -- every colour byte mapped from
// sRGB to linear with byte_color_to_linear (Image.cpp:40-51). A plugin built inside the reference tree would call IG::Image itself; this
// repository's build has no ig_runtime to link, so the one 8-bit format the reference's scenes use -- PNG -- is decoded here (own inflate,
// filters 0-4, grey / grey+alpha / RGB / RGBA / palette, 8 bits, non-interlaced), and so are float images in OpenEXR files (below).
#pragma once

#include <cstdint>
#include <string>
#include <vector>

namespace igbh {

struct DeviceImage {
    int format = 0;            // IGB200_IMAGE_RGBA8 | IGB200_IMAGE_MONO8 | IGB200_IMAGE_RGBA32F
    int width = 0, height = 0;
    std::vector<uint8_t> bytes;   // 8-bit formats, rows bottom-up
    std::vector<float> floats;    // RGBA32F (load_float_image), rows bottom-up
    const void* pixels() const { return floats.empty() ? static_cast<const void*>(bytes.data()) : static_cast<const void*>(floats.data()); }
    size_t pixel_bytes() const { return floats.empty() ? bytes.size() : floats.size() * sizeof(float); }
};

// Throws RecognizeError (script_recognizer.h) with the reason when the file cannot be read or is not a PNG this reader knows.
DeviceImage load_packed_image(const std::string& path, bool already_linear);

// Float images -- OpenEXR files (`device.load_image_by_id`: environment maps, the sky texture SkyLight.cpp writes into the cache). What the
// reference's device keeps (IG::Image::load, Image.cpp:500-712; Device.cpp:735-799): RGBA float, rows bottom-up, alpha 1 where the file has
// none, a single grey channel spread over R, G and B. Decoded here: single-part scanline files with HALF / FLOAT / UINT channels, compression
// NONE, RLE, ZIPS, ZIP and PIZ (own Huffman + wavelet decoder) -- what OpenEXR writers produce by default; tiled, deep, multi-part files and the
// lossy B44 / DWA / PXR24 compressions are reported. Radiance RGBE files (.hdr, flat or run-length encoded) are decoded as stb_image decodes them.
struct FloatImage {
    int width = 0, height = 0;
    std::vector<float> rgba;   // 4 floats per pixel, rows bottom-up
};
FloatImage load_float_image(const std::string& path);   // throws RecognizeError

// byte_color_to_linear (Image.cpp:40-51) for all 256 values
const uint8_t* srgb_byte_to_linear_byte();

}  // namespace igbh
