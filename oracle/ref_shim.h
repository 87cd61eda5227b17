// ref_shim.h -- force-included (-include) when oracle/build_ref.py compiles the UNMODIFIED reference CUDA
// extensions against the torch headers of this image.  The reference was written for torch 1.4, where
// AT_DISPATCH_FLOATING_TYPES accepted `tensor.type()` (an at::DeprecatedTypeProperties); torch 2.x only
// overloads ::detail::scalar_type for at::ScalarType.  This header restores the missing overload, nothing else.
// TEST INFRASTRUCTURE ONLY (see oracle/build_ref.py).
#pragma once
#ifdef __cplusplus
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties &t) { return t.scalarType(); }
}  // namespace detail
#endif
