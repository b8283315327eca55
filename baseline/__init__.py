"""baseline/ -- the REFERENCE ARM of bench.py and the reference-side half of the parity tests.

baseline/_ref/ (git-ignored, staged by baseline/stage_ref.py from /root/reference in the build container, travels to the
GPU box with the gpurun snapshot) holds the reference's own Python files, byte for byte.  baseline/ref_env.py imports
them UNMODIFIED; the compiled extensions they expect (`pointnet2_cuda`, `iou3d_cuda`, `roipool3d_cuda`) are served either
by the reference's own kernels (oracle/_ref/libpointnet2_ref.so, unmodified sources) or by the product (epnet_b200).
Nothing under epnet_b200/ imports this package.
"""
