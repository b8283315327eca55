// Stand-in for <ATen/cuda/CUDAContext.h>; nothing from it is used by the reference's .cu files.
#pragma once
