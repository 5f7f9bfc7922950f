#!/bin/bash
# Box probe: does the end-to-end figure depend on which NUMA node the host buffers live on?
nvidia-smi topo -m 2>/dev/null | head -12
cat /sys/bus/pci/devices/$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i 0 | tr 'A-Z' 'a-z' | sed 's/^0000//')/numa_node 2>/dev/null
for n in /sys/devices/system/node/node*; do
  cl=$(cat $n/cpulist); echo "== $(basename $n) cpus $cl"
  taskset -c $cl python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-binary --e2e-steps 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e G/s %.3f ms %.1f' % (d['e2e']['value']/1e9, d['e2e']['ms_per_step']))"
done
