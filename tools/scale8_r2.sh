mkdir -p gpurun_out
run() { # n tag envs...
  n=$1; tag=$2; shift 2
  env "$@" timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+RANDOM%400)) bench.py --gpus $n --no-cpu > gpurun_out/r2_scale_c2_$tag.json 2> gpurun_out/r2_scale_c2_$tag.err
  python - gpurun_out/r2_scale_c2_$tag.json <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split("/")[-1], "value %.4g e2e %.4g ms %.1f identical %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d.get("ranks_bit_identical_parameters")))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
}
run 8 n8 A=1
run 8 n8_direct FSRL_PPO_DP_DIRECT=1
run 4 n4 A=1
tail -3 gpurun_out/r2_scale_c2_n8.err
