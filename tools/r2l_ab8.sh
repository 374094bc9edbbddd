#!/bin/bash
# 8-GPU A/B of the exchange ordering protocol and the row transport (cfg2, bilinear headline + nearest).
mkdir -p gpurun_out
run() {  # N publish multicast
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $1 --steps 20 --warmup 3 --regions 7 --merge-publish $2 --multicast $3 2>gpurun_out/r2l_n$1_$2_$3_err.log > gpurun_out/r2l_n$1_$2_$3.json
  python - $1 $2 $3 <<'PY'
import json, sys
n, pub, mc = sys.argv[1:]
try:
    d = json.loads(open(f"gpurun_out/r2l_n{n}_{pub}_{mc}.json").read().strip().splitlines()[-1])
    v = {k: round(x["ms_per_step"] * 1e3, 2) for k, x in d["variants"].items()}
    print(f"N={n} publish={pub} multicast={mc}: {v} parity {d['parity']['ok']} e2e {d['e2e']['ms_per_step']:.3f} ms")
except Exception as e:
    print(f"N={n} publish={pub} multicast={mc}: FAILED {e}")
PY
}
run 8 kernel on
run 8 kernel off
run 8 host off
run 4 kernel on
run 2 kernel off
