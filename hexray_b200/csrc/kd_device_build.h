// kd_device_build.h — the KD-tree of a mesh built ON THE GPU (device/kdbuild.h), packed for the walk like a host-built one.
#pragma once
#include <string>
#include "device/launch.h"
#include "host/kdtree.h"

namespace hxr {

// false (err filled) if a device pass or an allocation failed; the context's sticky error is left for the caller to clear
bool buildKdTreeOnDevice(dev::Context* c, const hxr_mesh& mesh, const host::KdBuildParams& params, host::KdTree& out, std::string& err);

}  // namespace hxr
