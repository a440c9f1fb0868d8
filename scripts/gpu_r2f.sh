#!/bin/bash
# Round 2, visit F: full GPU suite, bench lines of every BASELINE config, ncu summaries of every kernel
# (the .ncu-rep files stay on the box: only the text summaries travel back)
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2f}
python scripts/dbg_corr140.py > ${O}_dbg_corr.log 2>&1; cat ${O}_dbg_corr.log | tail -12
timeout 1800 python -m pytest tests -m gpu -q --durations=6 > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 ${O}_pytest.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; tail -2 ${O}_bench_$name.err; }
if [ "$2" != "nobench" ]; then
b box16m
b jam1k --workload jam1k --steps 100000 --warmup 200
b jam65k --workload jam65k --steps 10000 --warmup 200 --no-cpu
b obs1m --workload obs1m --steps 10000 --warmup 200 --no-cpu
b sweep512 --workload sweep512 --steps 2000 --warmup 200 --no-cpu
b box16m_hot --workload box16m_hot --steps 200 --warmup 50 --no-cpu --no-e2e
b box2m --particles 2097152 --no-cpu --no-e2e
fi
if [ "$1" = "prof" ]; then
T=/tmp/ncu_$$; mkdir -p $T
cap() { out=$1; k=$2; skip=$3; shift 3; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 --kill 1 -o $T/$out -f "$@" > ${O}_ncu_$out.log 2>&1; rc=$?; python scripts/ncu_summary.py $T/$out.ncu-rep 14 > ${O}_${out}_ncu_full.txt 2>&1; echo "ncu $out rc=$rc $(head -1 ${O}_${out}_ncu_full.txt)"; rm -f $T/$out.ncu-rep; }
for k in apj_obs_reduce_kernel apj_obs_final_kernel apj_spatial_kernel apj_velhist_kernel apj_occupancy_kernel apj_scan_cells_kernel apj_scan_chunk_sums_kernel apj_scan_chunk_offsets_kernel apj_scatter_kernel apj_cell_sort_kernel apj_make_tiles_kernel apj_fill_tiles_kernel apj_finish_rebuild_kernel apj_reduce_commit_kernel apj_pack_kernel apj_unpack_kernel apj_checksum_kernel apj_mark_origin_kernel; do
  cap $k $k 1 python bench.py --workload obs1m --steps 200 --warmup 20 --no-relax --no-cpu
done
cap apj_step_kernel apj_step_kernel 150 python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16
cap apj_verlet_build_kernel apj_verlet_build_kernel 3 python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16
cap apj_reorder_kernel apj_reorder_kernel 3 python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16
cap apj_bin_count_kernel apj_bin_count_kernel 3 python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16
for k in apj_step_kernel apj_reduce_commit_kernel apj_push_ghosts_kernel apj_absorb_kernel apj_slab_sync_kernel apj_bin_count_kernel; do
  cap slab_$k $k 3 python scripts/slab_one_gpu.py 2097152 2 120
done
rm -rf $T
fi
