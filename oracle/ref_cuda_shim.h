/*
 * oracle/ref_cuda_shim.h -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
 *
 * Force-included (nvcc -include) when the reference's own CUDA extension sources are compiled
 * *where they lie* under /root/reference/FourierGrid/cuda for the GPU oracle (oracle/_ref/).
 * The reference targets torch 1.13 and writes AT_DISPATCH_FLOATING_TYPES(x.type(), ...); on
 * torch 2.11 DeprecatedTypeProperties no longer converts to ScalarType, so the unmodified
 * sources do not compile.  Instead of patching (or copying) the sources, this shim re-defines
 * the dispatch macro so that both spellings are accepted.  Nothing else is touched.
 */
#pragma once
#include <torch/extension.h>

namespace ubn_ref_shim {
inline at::ScalarType to_scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
inline at::ScalarType to_scalar_type(at::ScalarType t) { return t; }
}  // namespace ubn_ref_shim

#undef AT_DISPATCH_FLOATING_TYPES
#define AT_DISPATCH_FLOATING_TYPES(TYPE, NAME, ...) \
  AT_DISPATCH_SWITCH(::ubn_ref_shim::to_scalar_type(TYPE), NAME, AT_DISPATCH_CASE_FLOATING_TYPES(__VA_ARGS__))
