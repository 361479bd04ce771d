#!/bin/bash
# 1-GPU final pass: the whole gpu suite, smoke(), the driver's bench command, the reference arm, the other configs
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_pytest_final.log
tail -3 gpurun_out/r2_pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_reference_arm.json 2> gpurun_out/r2_bench_c3_reference_arm.err; echo "ref exit $?"
for wl in c1 c2 c4; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_$wl.json 2> gpurun_out/r2_bench_$wl.err; echo "bench $wl exit $?"; done
python - <<'PY'
import json
for f in ['c3_n1','c1','c2','c4']:
    b=json.load(open(f'gpurun_out/r2_bench_{f}.json'))
    print(f,'value',round(b['value']),'ms/step',round(b['ms_per_step'],3),'single',round(b['single_frame']['ms_per_step'],3),'one lane',round(b['single_frame']['ms_per_step_one_lane'],3),'fpb1',b['frames_per_batch_1'] and round(b['frames_per_batch_1']['ms_per_step'],3),'e2e',round(b['e2e']['value']),round(b['e2e']['ms_per_step'],3),'frac',round(b['roofline']['frac'],3),'dram_frac',b['roofline']['dram_frac'], b.get('image_check',{}).get('bit_identical'), b['clocks']['sm_mhz'], b['clocks']['reasons'])
r=json.load(open('gpurun_out/r2_bench_c3_reference_arm.json')); print('reference', r['value'], r['cpu_baseline'], r['product_lib_loaded'])
PY
