// oracle/shim/shim_globals.cpp -- TEST INFRASTRUCTURE, not product code.
// Storage for the glog FLAGS_* stand-ins (see glog/logging.h in this
// directory) and stubs for the two Ptex texture factories api.cpp:642,678
// refers to (textures/ptex.cpp needs the Ptex library; no config uses Ptex).
#include <glog/logging.h>

#include "pbrt.h"
#include "paramset.h"
#include "texture.h"
#include "transform.h"
#include "textures/ptex.h"

int FLAGS_stderrthreshold = 1;
int FLAGS_minloglevel = 0;
int FLAGS_v = 0;
bool FLAGS_logtostderr = false;
std::string FLAGS_log_dir;

namespace pbrt {
PtexTexture<Float> *CreatePtexFloatTexture(const Transform &, const TextureParams &) {
    Error("Ptex textures are not built into the oracle reference");
    return nullptr;
}
PtexTexture<Spectrum> *CreatePtexSpectrumTexture(const Transform &, const TextureParams &) {
    Error("Ptex textures are not built into the oracle reference");
    return nullptr;
}
}  // namespace pbrt
