set -x
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 5 > gpurun_out/r02_bench_n8_$tag.json 2> gpurun_out/r02_bench_n8_$tag.err; python - <<PY
import json
try:
    j=json.loads([l for l in open('gpurun_out/r02_bench_n8_$tag.json') if l.startswith('{')][-1])
    print("$tag N=8 value %.1f M/s ms %.2f map_ms %.2f e2e %.1f M/s" % (j['value']/1e6, j['ms_per_step'], j['mapping_stream_ms_per_step'], j['e2e']['value']/1e6), j['e2e']['per_step_ms'], j['kernel_ms'])
except Exception as e: print("$tag ERR", e)
PY
}
run p X=1
