#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/${TAG:-r2g}
timeout 1800 python -m pytest tests -m gpu -q --durations=6 > ${O}_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 ${O}_pytest.log
b() { name=$1; shift; timeout 900 python bench.py "$@" > ${O}_bench_$name.json 2> ${O}_bench_$name.err; echo "bench $name rc=$?"; python scripts/bench_brief.py ${O}_bench_$name.json; tail -2 ${O}_bench_$name.err; }
b obs1m --workload obs1m --steps 10000 --warmup 200 --no-cpu
b box16m --no-cpu
T=/tmp/ncu_$$; mkdir -p $T
cap() { out=$1; k=$2; skip=$3; shift 3; timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 --kill 1 -o $T/$out -f "$@" > ${O}_ncu_$out.log 2>&1; rc=$?; python scripts/ncu_summary.py $T/$out.ncu-rep 14 > ${O}_${out}_ncu_full.txt 2>&1; echo "ncu $out rc=$rc $(head -1 ${O}_${out}_ncu_full.txt)"; rm -f $T/$out.ncu-rep; }
cap apj_spatial_kernel apj_spatial_kernel 1 python bench.py --workload obs1m --steps 200 --warmup 20 --no-relax --no-cpu
cap slab_apj_step_kernel apj_step_kernel 60 python scripts/slab_one_gpu.py 2097152 2 120
cap apj_step_kernel apj_step_kernel 150 python bench.py --no-relax --no-cpu --no-e2e --steps 64 --warmup 16
cap apj_overlap_hue_kernel apj_overlap_hue_kernel 0 python -c "
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from bench import synthetic_state
from active_particle_jamming_b200 import DeviceEngine
R,L,x,y,phi=synthetic_state(1048576,1.0,3)
e=DeviceEngine(1048576,L,max_neighbors=64); e.upload(x=x,y=y,R=R,phi=phi); e.set_activity(0.05,0.5); e.skip_self_term_once(); e.step(200); print(e.overlap_hue()[:8]); e.init_lattice(5); print(e.box_table()[0][:2])"
cap apj_init_lattice_kernel apj_init_lattice_kernel 0 python -c "
import sys; sys.path.insert(0,'.')
from active_particle_jamming_b200 import DeviceEngine
L=DeviceEngine.lattice_box_length(1048576,0.9,5)[0]
e=DeviceEngine(1048576,L); e.init_lattice(5); print(e.counters())"
rm -rf $T
