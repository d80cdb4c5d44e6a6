/* oracle/ref_correlation_shim.h — TEST INFRASTRUCTURE.  Force-included (-include) when oracle/build_ref_correlation.sh compiles
 * the reference's unmodified correlation extension against a current PyTorch: AT_DISPATCH_* no longer accepts `tensor.type()`
 * (at::DeprecatedTypeProperties); this restores the overload it used to resolve to.  Nothing else is changed. */
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>
namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
