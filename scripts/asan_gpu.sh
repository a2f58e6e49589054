#!/bin/bash
# AddressSanitizer pass over the host side of the library (planner, evaluator, spill queue, codec):
# build an instrumented copy next to the normal one and run GPU tests against it.
#   scripts/asan_gpu.sh build            here (no GPU): kanter_core_b200/build/asan/libkanter_b200.so
#   scripts/asan_gpu.sh run [pytest args]   on the GPU box, e.g. through gpurun
set -e
cd "$(dirname "$0")/.."
OUT=kanter_core_b200/build/asan
ASAN=$(gcc -print-file-name=libasan.so)
if [ "$1" = build ]; then
    python -c "import sys; sys.path.insert(0, 'kanter_core_b200'); import build; build.write_jit_prelude()"
    mkdir -p $OUT
    for s in kc_context kc_kernels kc_fusion kc_h2n kc_resize kc_graph kc_exec kc_png kc_jit kc_numa; do
        nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O1 -g -std=c++17 --fmad=false \
            -Xcompiler -fPIC,-fsanitize=address,-fno-omit-frame-pointer -cudart static \
            -c kanter_core_b200/csrc/$s.cu -o $OUT/$s.o &
    done
    wait
    nvcc -shared -cudart static -Wno-deprecated-gpu-targets -o $OUT/libkanter_b200.so $OUT/*.o -lz -ldl -Xcompiler -fsanitize=address
    echo $OUT/libkanter_b200.so
else
    shift || true
    # protect_shadow_gap=0: the CUDA driver maps memory where ASan keeps its shadow gap.
    # log_path: pytest captures fd 2, and ASan leaves with _exit, so a report on stderr would be lost.
    mkdir -p gpurun_out
    KANTER_B200_LIB=$PWD/$OUT/libkanter_b200.so LD_PRELOAD=$ASAN \
        ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0:log_path=$PWD/gpurun_out/asan_report \
        python -u -m pytest -m gpu -v -p no:cacheprovider "$@"
fi
