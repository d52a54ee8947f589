lscpu | head -40 > gpurun_out/r02_host.txt 2>&1
(numactl -H || echo "no numactl") >> gpurun_out/r02_host.txt 2>&1
nvidia-smi topo -m >> gpurun_out/r02_host.txt 2>&1
ls /sys/devices/system/node/ >> gpurun_out/r02_host.txt 2>&1
cat /sys/devices/system/node/node*/cpulist >> gpurun_out/r02_host.txt 2>&1
nvidia-smi --query-gpu=index,pci.bus_id --format=csv >> gpurun_out/r02_host.txt 2>&1
for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d $(cat $d/numa_node) $(cat $d/class)"; fi; done >> gpurun_out/r02_host.txt 2>&1
grep -o -w -e avx2 -e avx512f -e avx512bw -e bmi2 /proc/cpuinfo | sort | uniq -c >> gpurun_out/r02_host.txt 2>&1
python /tmp/none.py 2>/dev/null
python - <<'PY' >> gpurun_out/r02_host.txt 2>&1
import time, numpy as np
a=np.random.rand(65536)
def t(f,n=30):
    f(); ts=[]
    for _ in range(n):
        t0=time.perf_counter(); f(); ts.append(time.perf_counter()-t0)
    return 1e6*np.median(ts)
print("tolist f64 65536 us", t(lambda:a.tolist()))
print("[0]*n us", t(lambda:[0]*65536))
import os; print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
PY
python bench.py --steps 50 --warmup 5 --sweep > gpurun_out/r02_bench_base.json 2> gpurun_out/r02_bench_base.err
