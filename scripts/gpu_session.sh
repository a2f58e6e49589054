#!/bin/bash
# One GPU-box session: parity tests, bench (both arms), per-kernel table, ncu launch list,
# ncu --set full captures of every kernel of the path.  Everything lands in gpurun_out/.
# usage: scripts/gpu_session.sh [tag] [skip-tests]     (tag defaults to r02)
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
if [ -z "$2" ]; then
python -m pytest tests -m gpu -q 2>&1 | tail -5 > $O/${TAG}_pytest.log
python bench.py > $O/${TAG}_bench.log 2>&1
python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref.log 2>&1
fi
python scripts/kernel_bench.py --math fast --out $O/kernels_${TAG}_fast.json > /dev/null 2> $O/${TAG}_kernels_fast.log
python scripts/kernel_bench.py --math exact --out $O/kernels_${TAG}_exact.json > /dev/null 2> $O/${TAG}_kernels_exact.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/launches_bench_${TAG}.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu > $O/${TAG}_ncu_launches.log 2>&1
# the reports are tens of MB each and gpurun_out/ only carries 64 MiB back: summarise on the box (ncu -i reads them here
# just as well), keep the summaries, drop the reports
summ() {  # name
  python scripts/ncu_summary.py $O/prof_$1_${TAG}.ncu-rep $O/ncu_full_$1_${TAG} > /dev/null 2>> $O/${TAG}_ncu_summary.log
  rm -f $O/prof_$1_${TAG}.ncu-rep
}
full() {  # name kernel-regex skip math only
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_$1_${TAG} \
      python scripts/kernel_bench.py --math $4 --only $5 --reps 2 > $O/${TAG}_ncu_$1.log 2>&1
  summ $1
}
full fused_cfg2_fast "kc_jit_entry" 1 fast config2_fused
full interpreter_cfg2_fast "kc_tile_vm" 1 fast config2_fused
full fused_cfg2_exact "kc_jit_entry" 1 exact config2_fused
full resize_tma_lanczos3_fast kc_resize_tma 3 fast resize_lanczos3_1024_to_8192_plane
full resize_tma_lanczos3_exact kc_resize_tma 3 exact resize_lanczos3_1024_to_8192_plane
full resize_tma_rgba_node_fast kc_resize_tma 3 fast resize_lanczos3_1024_to_8192_rgba
full resize_down_v "kc_resize_v_tma|kc_resize_v_march" 2 fast resize_lanczos3_8192
full resize_down_h kc_resize_h_tile 2 fast resize_lanczos3_8192
full h2n_fast kc_h2n_vec 3 fast height_to_normal
full h2n_exact kc_h2n_vec 3 exact height_to_normal
full to_u8_rgba "kc_tile_vm|kc_jit_entry" 3 fast to_u8_rgba
full from_u8 kc_from_u8 2 fast from_u8
full fill_segment "kc_tile_vm|kc_jit_entry" 3 fast fill_constant
# HeightToNormal strip mode, ring of one rank: the fused exchange (one launch) and round 1's publish kernel
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kc_h2n_vec -s 5 -c 1 -f -o $O/prof_h2n_exchange_${TAG} \
    python scripts/h2n_strips.py --steps 6 --halo peer > $O/${TAG}_ncu_h2n_exchange.log 2>&1
summ h2n_exchange
timeout 300 ncu --set full --clock-control none --import-source on -k regex:kc_halo_publish -s 5 -c 1 -f -o $O/prof_halo_publish_${TAG} \
    python scripts/h2n_strips.py --steps 6 --halo peer3 > $O/${TAG}_ncu_halo_publish.log 2>&1
summ halo_publish
ls $O/ncu_full_*_${TAG}.txt | wc -l; du -sh $O
tail -3 $O/${TAG}_pytest.log; tail -c 400 $O/${TAG}_bench.log; cat $O/${TAG}_kernels_fast.log
