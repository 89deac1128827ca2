// oracle/shim/imageio_pfm.cpp -- TEST INFRASTRUCTURE, not product code.
//
// Stand-in for the reference's core/imageio.cpp (which needs OpenEXR, lodepng
// and targa).  The oracle only ever writes float32 PFM images -- EXR output is
// half precision in the reference (core/imageio.cpp:164-189), useless for a
// 1e-4 parity check -- so this file implements the three functions declared in
// core/imageio.h:47-56 for ".pfm" only.  Row order and header follow the PFM
// convention the reference's own writer uses (bottom row first, little endian,
// scale -1).
#include "imageio.h"

#include <cstdio>
#include <cstring>
#include <vector>

#include "spectrum.h"

namespace pbrt {

static bool EndsWithPfm(const std::string &name) {
    return name.size() >= 4 && name.compare(name.size() - 4, 4, ".pfm") == 0;
}

void WriteImage(const std::string &name, const Float *rgb,
                const Bounds2i &outputBounds, const Point2i &totalResolution) {
    (void)totalResolution;
    std::string out = name;
    if (!EndsWithPfm(out)) {
        Warning("oracle image shim writes PFM only; writing \"%s.pfm\"", name.c_str());
        out += ".pfm";
    }
    Vector2i res = outputBounds.Diagonal();
    FILE *fp = std::fopen(out.c_str(), "wb");
    if (!fp) {
        Error("Unable to open output PFM file \"%s\"", out.c_str());
        return;
    }
    std::fprintf(fp, "PF\n%d %d\n-1.000000\n", res.x, res.y);
    std::vector<float> row(3 * (size_t)res.x);
    for (int y = res.y - 1; y >= 0; --y) {
        for (int x = 0; x < 3 * res.x; ++x) row[x] = (float)rgb[(size_t)y * 3 * res.x + x];
        std::fwrite(row.data(), sizeof(float), row.size(), fp);
    }
    std::fclose(fp);
}

static RGBSpectrum *ReadPfm(const std::string &name, int *w, int *h) {
    FILE *fp = std::fopen(name.c_str(), "rb");
    if (!fp) return nullptr;
    char tag[3] = {0, 0, 0};
    float scale = 0;
    if (std::fscanf(fp, "%2s %d %d %f", tag, w, h, &scale) != 4 || std::fgetc(fp) == EOF) {
        std::fclose(fp);
        return nullptr;
    }
    int nc = !std::strcmp(tag, "PF") ? 3 : (!std::strcmp(tag, "Pf") ? 1 : 0);
    if (!nc || scale >= 0) {  // big-endian files are not produced by anything here
        std::fclose(fp);
        return nullptr;
    }
    std::vector<float> data((size_t)nc * *w * *h);
    size_t got = std::fread(data.data(), sizeof(float), data.size(), fp);
    std::fclose(fp);
    if (got != data.size()) return nullptr;
    RGBSpectrum *img = new RGBSpectrum[(size_t)*w * *h];
    for (int y = 0; y < *h; ++y)
        for (int x = 0; x < *w; ++x) {
            const float *p = &data[(size_t)nc * ((size_t)(*h - 1 - y) * *w + x)];
            Float c[3] = {p[0], p[nc > 1 ? 1 : 0], p[nc > 1 ? 2 : 0]};
            img[(size_t)y * *w + x] = RGBSpectrum::FromRGB(c);
        }
    return img;
}

std::unique_ptr<RGBSpectrum[]> ReadImage(const std::string &name, Point2i *resolution) {
    if (EndsWithPfm(name)) {
        RGBSpectrum *img = ReadPfm(name, &resolution->x, &resolution->y);
        if (img) return std::unique_ptr<RGBSpectrum[]>(img);
    }
    Error("oracle image shim reads PFM only: \"%s\"", name.c_str());
    return nullptr;
}

RGBSpectrum *ReadImageEXR(const std::string &name, int *, int *, Bounds2i *, Bounds2i *) {
    Error("oracle image shim has no EXR reader: \"%s\"", name.c_str());
    return nullptr;
}

}  // namespace pbrt
