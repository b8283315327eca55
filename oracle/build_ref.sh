#!/bin/sh
# oracle/build_ref.sh -- TEST INFRASTRUCTURE.
# Compiles the reference's own pointnet2 kernels (the four *_gpu.cu files, UNMODIFIED, read where they
# lie under /root/reference) plus oracle/ref_shim.cu into oracle/_ref/libpointnet2_ref.so for sm_100a.
# Flags follow the reference's setup.py (nvcc -O2, default -fmad=true): that is what fixes the FMA
# contraction the index outputs depend on.  No reference source is copied into the repo; the .so is
# git-ignored but travels to the GPU box with the gpurun snapshot.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
SRC=${EPNET_REFERENCE_SRC:-/root/reference/pointnet2_lib/pointnet2/src}
OUT="$HERE/_ref"
if [ ! -d "$SRC" ]; then
    echo "build_ref.sh: $SRC not present (GPU box?) -- keeping prebuilt $OUT" >&2
    exit 0
fi
mkdir -p "$OUT"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-O2 -std=c++17 -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -I$HERE/stub -I$SRC"
for f in sampling_gpu ball_query_gpu group_points_gpu interpolate_gpu; do
    $NVCC $FLAGS -c "$SRC/$f.cu" -o "$OUT/$f.o"
done
ROI=${EPNET_REFERENCE_ROI:-/root/reference/lib/utils/roipool3d/src}
$NVCC $FLAGS -c "$ROI/roipool3d_kernel.cu" -o "$OUT/roipool3d_kernel.o"   # next-row oracle: the reference's RoI pooling kernels, unmodified
IOU=${EPNET_REFERENCE_IOU:-/root/reference/lib/utils/iou3d/src}
$NVCC $FLAGS -c "$IOU/iou3d_kernel.cu" -o "$OUT/iou3d_kernel.o"   # next-row oracle: the reference's rotated IoU / NMS kernels, unmodified
$NVCC $FLAGS -c "$HERE/ref_shim.cu" -o "$OUT/ref_shim.o"
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT/libpointnet2_ref.so" "$OUT"/*.o -lcudart
rm -f "$OUT"/*.o
echo "built $OUT/libpointnet2_ref.so"
