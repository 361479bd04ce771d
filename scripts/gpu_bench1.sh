#!/bin/bash
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_c3_reference_arm.json 2> gpurun_out/r2_bench_c3_reference_arm.err; echo "ref exit $?"
timeout 600 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_bench_c4.json 2> gpurun_out/r2_bench_c4.err; echo "c4 exit $?"
python - <<'PY'
import json
for f in ['c3_n1','c4']:
    b=json.load(open(f'gpurun_out/r2_bench_{f}.json')); rf=b['roofline']
    print(f,'value',round(b['value']),'ms/step',round(b['ms_per_step'],3),'single',round(b['single_frame']['ms_per_step'],3),'one lane',round(b['single_frame']['ms_per_step_one_lane'],3),'fpb1',b['frames_per_batch_1'] and round(b['frames_per_batch_1']['ms_per_step'],3),'e2e',round(b['e2e']['value']),round(b['e2e']['ms_per_step'],3),'frac',rf['frac'] and round(rf['frac'],3),'dram_frac',rf['dram_frac'], b.get('image_check',{}).get('bit_identical'), b['clocks']['sm_mhz'], b['clocks']['reasons'], b['config']['frames_per_batch'])
r=json.load(open('gpurun_out/r2_bench_c3_reference_arm.json')); print('reference', r['value'], r['cpu_baseline']['cores'], r['product_lib_loaded'])
PY
