#!/bin/bash
# 8-GPU box: host topology, the raw pinned-copy ceiling with 1/2/4/8 GPUs copying at once, then the all-config bench at N=8
mkdir -p gpurun_out
{ nvidia-smi topo -m; echo; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)|Thread|Core"; echo; (numactl -H 2>/dev/null || echo "numactl not installed"); echo; free -g | head -2; echo; for d in /sys/bus/pci/devices/*; do :; done; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.width.current --format=csv; } > gpurun_out/n8_topology.txt 2>&1
for n in 1 2 4 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n scratch/pcie_multi.py 2> gpurun_out/pcie_n$n.err | tail -1 > gpurun_out/pcie_n$n.json
  cat gpurun_out/pcie_n$n.json
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 5 2> gpurun_out/bench_n8.err | tail -1 > gpurun_out/bench_n8.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read())
print('N=8 headline', '%.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'link', d['e2e'].get('link_gbs'), d['e2e'].get('frac_of_link'), d['clocks'])
for k,c in d['configs'].items(): print(k, '%.4g'%c['value'], 'frac %.3f'%c['roofline']['frac'], 'e2e %.4g'%c['e2e']['value'], 'link %.1f frac %.2f'%(c['e2e']['link_gbs'], c['e2e']['frac_of_link']))
PY
tail -5 gpurun_out/bench_n8.err
