#!/bin/bash
# 8 GPUs: the driver's line, then literal-config variants (200 steps)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
run() { # name, extra env..., args
  name=$1; shift
  timeout 600 env "$@" > /dev/null 2>&1
}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port"
timeout 900 $TR 29821 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/bench_n8.err
timeout 300 $TR 29822 bench.py --gpus 8 --steps 200 --warmup 5 --secondary none --no-cpu-baseline > gpurun_out/bench_n8_s200.json 2> gpurun_out/bench_n8_s200.err; echo "rc=$?"
COSMOSLIK_B200_MULTICAST=1 timeout 300 $TR 29823 bench.py --gpus 8 --steps 200 --warmup 5 --secondary none --no-weak --no-cpu-baseline > gpurun_out/bench_n8_mc.json 2> gpurun_out/bench_n8_mc.err; echo "mc rc=$?"
for t in 32 128 256; do
CSL_SAB_THREADS=$t timeout 300 $TR 2983$((t/32)) bench.py --gpus 8 --steps 200 --warmup 5 --secondary none --no-weak --no-cpu-baseline > gpurun_out/bench_n8_sab$t.json 2> gpurun_out/bench_n8_sab$t.err; echo "sab$t rc=$?"
done
CSL_SEG_SMALL=0 timeout 300 $TR 29841 bench.py --gpus 8 --steps 200 --warmup 5 --secondary none --no-weak --no-cpu-baseline > gpurun_out/bench_n8_seg2.json 2> gpurun_out/bench_n8_seg2.err; echo "seg2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_n8*.json')):
    try: d=json.load(open(f))
    except Exception as e: print(f, 'unreadable', e); continue
    print(f, 'literal %.4g evals/s %.4f ms' % (d['value'], d['ms_per_step']), 'ranks', [round(x,4) for x in d['ms_per_rank']], d.get('exchange_parity'), d['exchange'][:60])
    if d.get('weak'): print('    weak %.4g %.4f' % (d['weak']['value'], d['weak']['ms_per_step']), [round(x,4) for x in d['weak']['ms_per_rank']])
    if d.get('roofline'): print('    ', {k:(round(v['ms_per_launch'],4), round(v['frac_of_fp64_peak'],3)) for k,v in d['roofline']['kernels'].items()}, d['roofline']['other_stages_ms'])
    for s in d.get('secondary') or []: print('    ', s['name'], s.get('error'), s.get('value'), s.get('ms_per_step'))
PY
