for m in 0 1 0 1; do
  FOVGS_NO_DIRECT_STATS=$m python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$m bench.py --gpus 2 --steps 40 --warmup 8 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('no_direct_stats=$m', 'value %.1f'%d['value'], 'e2e %.1f'%d['e2e']['value'], 'pipelined %.1f'%d['e2e']['pipelined_value'], 'u8 %.1f'%d['e2e']['uint8_output']['value'])
"
done
