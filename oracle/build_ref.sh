#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- never on the product path.
#
# Builds the reference's OWN native sources for the 3D RoI hot path, from where they lie under
# /root/reference, into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun):
#
#   oracle/_ref/ref_nms_cpu*.so        mmdet/ops/nms/src/nms_cpu.cpp            (CPU, runs anywhere)
#   oracle/_ref/ref_nms_cuda*.so       mmdet/ops/nms/src/{nms_cuda.cpp,nms_kernel.cu}       (needs a GPU)
#   oracle/_ref/ref_roi_align_cuda*.so mmdet/ops/roi_align/src/{roi_align_cuda.cpp,roi_align_kernel.cu}
#
# The reference targets PyTorch 1.0 / THC.  Its sources are copied to a scratch directory OUTSIDE
# the repo and patched there by the sed lines below, which only rename removed PyTorch-1.0 APIs
# (.type() -> .scalar_type(), THCudaCheck -> C10_CUDA_CHECK, THCudaMalloc -> caching allocator,
# AT_CHECK -> TORCH_CHECK ...).  No arithmetic line is touched, no nvcc math flag is added (the
# reference sets none: mmdet/ops/nms/setup.py:73-84, mmdet/ops/roi_align/setup.py:4-12), so IEEE
# division and nvcc's default -fmad=true contraction are exactly what the reference ships with.
# No reference source is copied into this repository.
set -euo pipefail

REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
[ -d "$REF/mmdet/ops" ] || { echo "build_ref: $REF not present; keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"
TMP=$(mktemp -d /tmp/roi3d_ref_build.XXXXXX)
trap 'rm -rf "$TMP"' EXIT

PY=${PYTHON:-python}
read -r TORCH_INC TORCH_INC2 TORCH_LIB PY_INC EXT < <($PY - <<'EOF'
import sysconfig, os, torch
base = os.path.dirname(torch.__file__)
print(os.path.join(base, "include"), os.path.join(base, "include/torch/csrc/api/include"),
      os.path.join(base, "lib"), sysconfig.get_paths()["include"], sysconfig.get_config_var("EXT_SUFFIX"))
EOF
)
CUDA_HOME=${CUDA_HOME:-/usr/local/cuda}
INC="-I$TORCH_INC -I$TORCH_INC2 -I$PY_INC -I$CUDA_HOME/include"
CXXFLAGS="-O2 -std=c++17 -fPIC -D_GLIBCXX_USE_CXX11_ABI=1 -DTORCH_API_INCLUDE_EXTENSION_H"
LIBS="-L$TORCH_LIB -Wl,-rpath,$TORCH_LIB -lc10 -ltorch_cpu -ltorch -ltorch_python"
CULIBS="$LIBS -lc10_cuda -ltorch_cuda -L$CUDA_HOME/lib64 -lcudart"
ARCH="-gencode arch=compute_100a,code=sm_100a"

# ---- CPU NMS (2-D semantics; the CPU baseline BASELINE.json config 1 names) -------------------
sed 's/AT_DISPATCH_FLOATING_TYPES(dets.type(), "nms"/AT_DISPATCH_FLOATING_TYPES(dets.scalar_type(), "nms"/;
     s/dets\.type()\.is_cuda()/dets.is_cuda()/g;
     s/\.data<\([a-z0-9_]*\)>()/.data_ptr<\1>()/g' \
    "$REF/mmdet/ops/nms/src/nms_cpu.cpp" > "$TMP/nms_cpu.cpp"
g++ $CXXFLAGS $INC -DTORCH_EXTENSION_NAME=ref_nms_cpu -shared "$TMP/nms_cpu.cpp" -o "$OUT/ref_nms_cpu$EXT" $LIBS

# ---- CUDA NMS (2-D + 3-D) ----------------------------------------------------------------------
sed -e 's|#include <THC/THC.h>|#include <c10/cuda/CUDACachingAllocator.h>\n#include <c10/cuda/CUDAException.h>|' \
    -e 's|#include <THC/THCDeviceUtils.cuh>|#include <ATen/ceil_div.h>|; s/THCCeilDiv/at::ceil_div/g' \
    -e 's/THCState \*state = .*$//' \
    -e 's/THCudaMalloc(state, /c10::cuda::CUDACachingAllocator::raw_alloc(/' \
    -e 's/THCudaFree(state, mask_dev)/c10::cuda::CUDACachingAllocator::raw_delete(mask_dev)/' \
    -e 's/THCudaCheck(/C10_CUDA_CHECK(/g; s/boxes\.type()\.is_cuda()/boxes.is_cuda()/g' \
    -e 's/\.data<scalar_t>()/.data_ptr<scalar_t>()/g; s/\.data<int64_t>()/.data_ptr<int64_t>()/g' \
    "$REF/mmdet/ops/nms/src/nms_kernel.cu" > "$TMP/nms_kernel.cu"
sed -e 's/AT_CHECK(/TORCH_CHECK(/g; s/x\.type()\.is_cuda()/x.is_cuda()/g' \
    "$REF/mmdet/ops/nms/src/nms_cuda.cpp" > "$TMP/nms_cuda.cpp"
nvcc $ARCH -O2 -std=c++17 -lineinfo -Xcompiler -fPIC -D_GLIBCXX_USE_CXX11_ABI=1 $INC \
    -c "$TMP/nms_kernel.cu" -o "$TMP/nms_kernel.o"
g++ $CXXFLAGS $INC -DTORCH_EXTENSION_NAME=ref_nms_cuda -c "$TMP/nms_cuda.cpp" -o "$TMP/nms_cuda.o"
g++ -shared "$TMP/nms_cuda.o" "$TMP/nms_kernel.o" -o "$OUT/ref_nms_cuda$EXT" $CULIBS

# ---- CUDA RoIAlign (2-D + 3-D, forward + backward) ---------------------------------------------
sed -e 's/features\.type()/features.scalar_type()/g; s/top_grad\.type()/top_grad.scalar_type()/g' \
    -e 's/THCudaCheck(/C10_CUDA_CHECK(/g; s/\.data<scalar_t>()/.data_ptr<scalar_t>()/g' \
    -e 's|#include <THC/THCAtomics.cuh>|#include <ATen/cuda/Atomic.cuh>\n#include <c10/cuda/CUDAException.h>|' \
    "$REF/mmdet/ops/roi_align/src/roi_align_kernel.cu" > "$TMP/roi_align_kernel.cu"
sed -e 's/AT_CHECK(/TORCH_CHECK(/g; s/x\.type()\.is_cuda()/x.is_cuda()/g' \
    "$REF/mmdet/ops/roi_align/src/roi_align_cuda.cpp" > "$TMP/roi_align_cuda.cpp"
nvcc $ARCH -O2 -std=c++17 -lineinfo -Xcompiler -fPIC -D_GLIBCXX_USE_CXX11_ABI=1 $INC \
    -c "$TMP/roi_align_kernel.cu" -o "$TMP/roi_align_kernel.o"
g++ $CXXFLAGS $INC -DTORCH_EXTENSION_NAME=ref_roi_align_cuda -c "$TMP/roi_align_cuda.cpp" -o "$TMP/roi_align_cuda.o"
g++ -shared "$TMP/roi_align_cuda.o" "$TMP/roi_align_kernel.o" -o "$OUT/ref_roi_align_cuda$EXT" $CULIBS

# keep the SASS of the reference kernels beside the binaries: it is the evidence for which
# mul+add pairs nvcc contracted to FFMA (the oracle restates exactly those).
cuobjdump -sass "$TMP/nms_kernel.o" > "$OUT/ref_nms_kernel.sass" 2>/dev/null || true
cuobjdump -sass "$TMP/roi_align_kernel.o" > "$OUT/ref_roi_align_kernel.sass" 2>/dev/null || true
echo "build_ref: wrote $(ls "$OUT" | tr '\n' ' ')"
