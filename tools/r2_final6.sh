for i in 1 2; do
AGS_E2E_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2962$i bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2final6_n2_run$i.json 2> gpurun_out/r2final6_n2_run$i.err
echo "== run $i"; grep "host trace" gpurun_out/r2final6_n2_run$i.err | cut -c1-1400
python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2final6_n2_run$i.json') if l.startswith('{')][-1]); print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], d['e2e']['host_ms_per_update'])"
done
