i=0
for v in "--e2e-threads 3 --blocking-sync" "--e2e-threads 3" "--e2e-threads 4 --no-numa-bind" "--e2e-threads 2" "--e2e-threads 6 --blocking-sync"; do
  i=$((i+1))
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29530+i)) bench.py --gpus 8 --steps 3 --warmup 3 --no-other-configs $v 2>gpurun_out/b8.err | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', round(d['value']), round(d['e2e']['value']), d['e2e']['threads'], d['e2e']['host'])"
done
