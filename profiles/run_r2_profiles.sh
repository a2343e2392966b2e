#!/bin/bash
# Round-2 ncu evidence, one B200: (1) the plain bench (must exit 0 first), (2) the launch list of the same command,
# (3) one --set full capture of the neighbor kernel and of the row kernel.  Outputs -> gpurun_out/.
TAG=${1:-r2z}
set -x
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"neighbors_kernel|rows_acc_kernel" -c 2 -s 6 -o gpurun_out/${TAG}_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_*
