#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), per-kernel table, ncu launch list,
# ncu --set full captures of the top kernels.  Everything lands in gpurun_out/.
# usage: scripts/gpu_session.sh [tag]     (tag defaults to r01)
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_bench_ref.log 2>&1
python scripts/kernel_bench.py --math fast --out $O/kernels_${TAG}_fast.json > /dev/null 2> $O/${TAG}_kernels_fast.log
python scripts/kernel_bench.py --math exact --out $O/kernels_${TAG}_exact.json > /dev/null 2> $O/${TAG}_kernels_exact.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_${TAG}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_launches.log 2>&1
full() {  # name kernel-regex skip math only
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_$1_${TAG} \
      python scripts/kernel_bench.py --math $4 --only $5 --reps 2 > $O/${TAG}_ncu_$1.log 2>&1
}
full tilevm_cfg2_fast "kc_tile_vm|kc_jit_entry" 3 fast config2_fused
full resize_lanczos3_fast kc_resize_strip 3 fast resize_lanczos3_1024
full h2n_fast kc_h2n_vec 3 fast height_to_normal
full h2n_exact kc_h2n_vec 3 exact height_to_normal
full to_u8_rgba "kc_tile_vm|kc_jit_entry" 3 fast to_u8_rgba
full from_u8 kc_from_u8 2 fast from_u8
tail -3 $O/${TAG}_pytest.log; tail -c 600 $O/${TAG}_bench.log; cat $O/${TAG}_kernels_fast.log
