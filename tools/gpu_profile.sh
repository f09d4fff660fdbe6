#!/bin/bash
# usage (on the GPU box, under gpurun): tools/gpu_profile.sh <tag>
# launch list of the bench command + ncu --set full of every kernel of one step (raw page as csv) + source-level
# captures of the luma search kernels; keeps gpurun_out/ under the 64 MiB that travel back.
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py --steps 2 --warmup 1 --no-cpu --no-photo --no-other > $O/bench_quick_$TAG.json 2> $O/bench_quick_$TAG.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-photo --no-other > $O/ncu_launch.log 2>&1
echo "launch list rc=$?"
python tools/gpu_prof_workload.py 1024 synthetic 1 || exit 1
timeout 600 ncu --set full --clock-control none -f -o $O/prof_${TAG}_all python tools/gpu_prof_workload.py 1024 synthetic 1 > $O/ncu_all.log 2>&1
echo "all rc=$?"
ncu -i $O/prof_${TAG}_all.ncu-rep --page raw --csv > $O/prof_${TAG}_all_raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_search -f -o $O/prof_${TAG}_search python tools/gpu_prof_workload.py 1024 synthetic 1 > $O/ncu_search.log 2>&1
echo "search rc=$?"
ZW_QUAD=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_searchq<\(int\)2>|k_searchq<2>" -f -o $O/prof_${TAG}_quad2 python tools/gpu_prof_workload.py 1024 synthetic 1 > $O/ncu_quad2.log 2>&1
echo "quad2 rc=$?"
for r in search quad2; do
  ncu -i $O/prof_${TAG}_$r.ncu-rep --page source --csv > $O/prof_${TAG}_${r}_source.csv 2>/dev/null
  ncu -i $O/prof_${TAG}_$r.ncu-rep --page raw --csv > $O/prof_${TAG}_${r}_raw.csv 2>/dev/null
done
gzip -f $O/*_source.csv
ls -la $O
# drop the largest reports until the directory fits
while [ $(du -sm $O | cut -f1) -ge 60 ]; do
  big=$(ls -S $O/*.ncu-rep 2>/dev/null | head -1)
  [ -z "$big" ] && break
  echo "dropping $big"; rm -f "$big"
done
du -sm $O
