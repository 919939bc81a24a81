#!/bin/bash
# c2 weak scaling at N = 8 and N = 4 (run under gpurun --gpus 8); N = 1 / 2 lines come from the 1- and 2-GPU calls
mkdir -p gpurun_out
run() {
  n=$1
  timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench.py --gpus $n --no-cpu > gpurun_out/r2_scale_c2_n$n.json 2> gpurun_out/r2_scale_c2_n$n.err
  python - gpurun_out/r2_scale_c2_n$n.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "value %.4g e2e %.4g ms %.1f identical %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("ranks_bit_identical_parameters")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
run 8
run 4
