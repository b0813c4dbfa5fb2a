#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (see oracle/README.md).
# Builds the UNMODIFIED reference CUDA extension (svox2/csrc/*.cu + svox2.cpp) for sm_100a from the
# sources where they lie under /root/reference into oracle/_ref/svox2_ref_csrc*.so.  Nothing is copied
# into this repository; oracle/_ref/ is git-ignored but travels to the GPU box with the snapshot.
# It is the GPU-side comparator for the `-m gpu` parity tests (reference kernels vs ours on the same B200).
# Torch's default CUDAExtension flags are mirrored (no fast-math), plus -DNDEBUG so device asserts
# (ASSERT_NUM) cannot trap the context on degenerate rays; arithmetic is unaffected.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=$OUT/obj
mkdir -p "$OBJ"
if [ ! -d "$REF/svox2/csrc" ]; then echo "reference not present, keeping prebuilt files"; exit 0; fi
PY=${PYTHON:-python}
TORCH_INC=$($PY -c "import torch.utils.cpp_extension as c; print(' '.join('-I'+p for p in c.include_paths()))" 2>/dev/null)
TORCH_LIB=$($PY -c "import torch, os; print(os.path.join(os.path.dirname(torch.__file__), 'lib'))" 2>/dev/null)
PY_INC=$($PY -c "import sysconfig; print(sysconfig.get_paths()['include'])")
SUFFIX=$($PY -c "import sysconfig; print(sysconfig.get_config_var('EXT_SUFFIX'))")
NAME=svox2_ref_csrc
COMMON="-I$REF/svox2/csrc/include $TORCH_INC -I$PY_INC -I/usr/local/cuda/include -DTORCH_EXTENSION_NAME=$NAME -DTORCH_API_INCLUDE_EXTENSION_H -DNDEBUG -D_GLIBCXX_USE_CXX11_ABI=1"
NVCC_FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC -w \
 -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_BFLOAT16_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__"
SRCS="svox2_kernel render_lerp_kernel_cuvol render_lerp_kernel_surface render_lerp_kernel_surf_trav render_lerp_kernel_nvol render_svox1_kernel misc_kernel loss_kernel optim_kernel test_cuda"
pids=()
for s in $SRCS; do
  if [ ! -f "$OBJ/$s.o" ] || [ "$REF/svox2/csrc/$s.cu" -nt "$OBJ/$s.o" ]; then
    ( nvcc $NVCC_FLAGS $COMMON -c "$REF/svox2/csrc/$s.cu" -o "$OBJ/$s.o" > "$OBJ/$s.log" 2>&1 || { echo "FAILED $s"; tail -20 "$OBJ/$s.log"; exit 1; } ) &
    pids+=($!)
  fi
done
if [ ! -f "$OBJ/svox2.o" ]; then
  ( g++ -O2 -std=c++17 -fPIC -w $COMMON -c "$REF/svox2/csrc/svox2.cpp" -o "$OBJ/svox2.o" > "$OBJ/svox2.log" 2>&1 || { echo "FAILED svox2.cpp"; tail -20 "$OBJ/svox2.log"; exit 1; } ) &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
OBJS=""; for s in $SRCS svox2; do OBJS="$OBJS $OBJ/$s.o"; done
g++ -shared -o "$OUT/$NAME$SUFFIX" $OBJS -L"$TORCH_LIB" -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python \
   -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$TORCH_LIB"
echo "built $OUT/$NAME$SUFFIX"
# The reference's own Python package, staged UNMODIFIED beside the comparator (git-ignored like the .so, travels to the GPU
# box): tests/test_dropin_gpu.py imports it twice -- with `svox2.csrc` = our compiled shim and = the reference extension --
# and bench.py's cpu_baseline times its pure-PyTorch gradcheck renderer (config C1) on the box's host cores.
mkdir -p "$OUT/pyref/svox2"
cp -f "$REF"/svox2/*.py "$OUT/pyref/svox2/"
echo "staged $OUT/pyref/svox2 ($(ls "$OUT/pyref/svox2" | wc -l) files)"
