#!/bin/bash
# Usage (on the GPU box): profiles/run_d2h_probe.sh <ranks> <tag>
# Host / PCIe facts + the D2H probe; output under gpurun_out/.
N=${1:-4}; TAG=${2:-r2}
OUT=gpurun_out/d2h_probe_${TAG}_n${N}
mkdir -p gpurun_out
{
  echo "== lscpu"; lscpu | egrep -i 'model name|socket|numa|^cpu\(s\)|thread|core|hypervisor|virtual|L3'
  echo "== nproc $(nproc)"; echo "== affinity"; taskset -p $$ 
  echo "== numa nodes"; ls /sys/devices/system/node/ | tr '\n' ' '; echo
  for n in /sys/devices/system/node/node*; do echo "$n $(cat $n/cpulist) $(grep MemTotal $n/meminfo)"; done
  echo "== thp"; cat /sys/kernel/mm/transparent_hugepage/enabled
  echo "== uncore"; ls /sys/devices/system/cpu/intel_uncore_frequency 2>/dev/null | head; cat /sys/devices/system/cpu/cpu0/cpufreq/scaling_governor 2>/dev/null
  echo "== cpuidle"; cat /sys/devices/system/cpu/cpuidle/current_driver 2>/dev/null; ls /sys/devices/system/cpu/cpu0/cpuidle 2>/dev/null | tr '\n' ' '; echo
  echo "== topo"; nvidia-smi topo -m
  echo "== pcie"; nvidia-smi --query-gpu=index,pci.bus_id,pcie.link.gen.current,pcie.link.gen.max,pcie.link.width.current,pcie.link.width.max --format=csv
  echo "== gpu numa"; for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$d numa=$(cat $d/numa_node) class=$(cat $d/class)"; fi; done
  echo "== free"; free -g | head -2
} > ${OUT}_host.txt 2>&1
timeout 240 profiles/_bin/d2h_probe $N 4 > ${OUT}.jsonl 2> ${OUT}.err
echo "probe rc=$?"
tail -c 600 ${OUT}.err
