// Stand-in for <torch/serialize/tensor.h>: the reference's *_gpu.h headers only need at::Tensor to
// exist as a name (used by-value in wrapper DECLARATIONS that the oracle build never defines).
#pragma once
namespace at { class Tensor; }
