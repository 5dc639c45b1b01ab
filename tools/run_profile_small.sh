#!/bin/bash
# Dev helper (GPU box): launch list + set-full capture of the lane kernels of one small workload, summarised ON THE BOX (the reports
# exceed the 64 MiB that travel back).  usage: bash tools/run_profile_small.sh <workload> <profiles tag> "<description>"
w=$1; tag=$2; desc=$3
python tools/one_step.py $w 3 > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv python tools/one_step.py $w 3 > gpurun_out/${tag}_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:ncde_lane -c 4 -o /tmp/${tag}_full python tools/one_step.py $w 1 > gpurun_out/${tag}_ncu2.log 2>&1; tail -1 gpurun_out/${tag}_ncu2.log
python tools/make_profiles.py $tag gpurun_out/${tag}_launches.csv /tmp/${tag}_full.ncu-rep "python tools/one_step.py $w 3 / 1   ($desc)" 2>&1 | tail -1
cp profiles/${tag}_* gpurun_out/ 2>/dev/null
