#!/bin/bash
# Round-2 measurement pass (run under gpurun).  Usage: tools/bench_all_r2.sh [single|multi]
mode=${1:-single}
mkdir -p gpurun_out
if [ "$mode" = "single" ]; then
  timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_c2_ref.json 2> gpurun_out/r2_bench_c2_ref.err
  timeout 600 python bench.py > gpurun_out/r2_bench_c2_n1.json 2> gpurun_out/r2_bench_c2_n1.err
  timeout 300 python bench.py --config c1 --impl reference > gpurun_out/r2_bench_c1_ref.json 2> gpurun_out/r2_bench_c1_ref.err
  timeout 300 python bench.py --config c1 > gpurun_out/r2_bench_c1_n1.json 2> gpurun_out/r2_bench_c1_n1.err
  timeout 600 python bench.py --config c5 --steps 2 --warmup 1 > gpurun_out/r2_bench_c5_n1.json 2> gpurun_out/r2_bench_c5_n1.err
  timeout 900 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_bench_c4_n1.json 2> gpurun_out/r2_bench_c4_n1.err
  timeout 900 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/r2_bench_c3_n1.json 2> gpurun_out/r2_bench_c3_n1.err
else
  for n in 1 2 4 8; do
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --no-cpu > gpurun_out/r2_scale_c2_n$n.json 2> gpurun_out/r2_scale_c2_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) \
        bench.py --gpus $n --no-cpu > gpurun_out/r2_scale_c2_n$n.json 2> gpurun_out/r2_scale_c2_n$n.err
    fi
  done
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29610 \
    bench.py --config c5 --gpus 8 --steps 2 --warmup 1 --no-cpu > gpurun_out/r2_scale_c5_n8.json 2> gpurun_out/r2_scale_c5_n8.err
fi
for f in gpurun_out/r2_*${mode:0:1}*.json gpurun_out/r2_bench_*.json gpurun_out/r2_scale_*.json; do
  [ -s "$f" ] && python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "value %.4g" % d.get("value", float("nan")), "e2e %.4g" % d.get("e2e", {}).get("value", float("nan")), "ms/step %.1f" % d.get("ms_per_step", float("nan")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
