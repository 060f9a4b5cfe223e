// ORACLE tooling.  Force-included (-include) when compiling the UNMODIFIED reference source
// /root/reference/pysgg/csrc/cpu/ROIAlign_cpu.cpp against torch >= 2: the only incompatibility is
// AT_DISPATCH_FLOATING_TYPES(input.type(), ...) receiving a DeprecatedTypeProperties (ROIAlign_cpu.cpp:242).
#pragma once
#include <torch/extension.h>
#include <ATen/Dispatch.h>
namespace veto_compat {
inline at::ScalarType st(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
inline at::ScalarType st(at::ScalarType t) { return t; }
}  // namespace veto_compat
#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(veto_compat::st(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
