#!/bin/bash
# Environment probe of the GPU box: managed runtimes (a real NVorbis run would pin parity), NUMA / PCIe topology (pinned-buffer placement).
mkdir -p gpurun_out
{
  echo "== managed runtimes"; for t in dotnet mono csc mcs msbuild xbuild; do printf "%s: " $t; (command -v $t || echo absent); done
  dotnet --version 2>&1 | head -2; mono --version 2>&1 | head -2
  echo "== cpu"; nproc; lscpu | grep -i -E "model name|socket|numa|thread|core" 
  echo "== numa"; (command -v numactl && numactl -H) 2>&1 | head -20; ls /sys/devices/system/node/ 2>&1 | head; cat /sys/devices/system/node/node*/cpulist 2>&1
  echo "== topo"; nvidia-smi topo -m 2>&1 | head -40
  echo "== gpus"; nvidia-smi --query-gpu=index,name,pci.bus_id,clocks.sm,clocks.max.sm,power.draw --format=csv
  for d in /sys/bus/pci/devices/*; do if [ -f $d/numa_node ] && grep -q 0x10de $d/vendor 2>/dev/null; then echo "$d numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist)"; fi; done
  echo "== affinity"; taskset -p $$; grep -i cpus_allowed_list /proc/self/status; grep -i mems_allowed_list /proc/self/status
  echo "== libnuma"; ls /usr/lib/x86_64-linux-gnu/libnuma* 2>&1; python -c "import ctypes; print(ctypes.CDLL('libnuma.so.1'))" 2>&1
  echo "== libavcodec"; python -c "import sys; sys.path.insert(0,'tests'); import ffmpeg_vorbis as F; print('ffmpeg cross-check available:', F.available() if hasattr(F,'available') else 'n/a')" 2>&1 | tail -2
  echo "== mem"; free -g | head -3
} > gpurun_out/env_probe.txt 2>&1
cat gpurun_out/env_probe.txt
