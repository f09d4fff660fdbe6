#!/bin/bash
# usage (on the GPU box, under gpurun): tools/gpu_profile.sh <tag>
# launch list of the bench command + ncu --set full of every kernel of one step (raw page as csv), the same for the photo
# workload and the lossless path, + a source-level capture of the luma search kernels; keeps gpurun_out/ small.
TAG=${1:-r2}
O=gpurun_out
mkdir -p $O
python bench.py --steps 2 --warmup 1 --no-cpu --no-photo --no-other --no-verify > $O/bench_quick_$TAG.json 2> $O/bench_quick_$TAG.err || exit 1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --no-photo --no-other --no-verify > $O/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_photo_$TAG.csv \
  python tools/gpu_prof_workload.py 1024 photo 1 > $O/ncu_launch_photo.log 2>&1
echo "photo launch list rc=$?"
python tools/gpu_prof_workload.py 1024 synthetic 1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -f -o $O/prof_${TAG}_all python tools/gpu_prof_workload.py 1024 synthetic 1 > $O/ncu_all.log 2>&1
echo "all rc=$?"
ncu -i $O/prof_${TAG}_all.ncu-rep --page raw --csv > $O/prof_${TAG}_all_raw.csv 2>/dev/null
ncu -i $O/prof_${TAG}_all.ncu-rep --page source --csv > $O/prof_${TAG}_all_source.csv 2>/dev/null
if [ -z "$SKIP_LOSSLESS" ]; then  # SKIP_LOSSLESS=1: the lossless kernels did not change since the last capture
timeout 600 ncu --set full --clock-control none -f -o $O/prof_${TAG}_lossless python tools/gpu_prof_lossless.py 1024 > $O/ncu_lossless.log 2>&1
echo "lossless rc=$?"
ncu -i $O/prof_${TAG}_lossless.ncu-rep --page raw --csv > $O/prof_${TAG}_lossless_raw.csv 2>/dev/null
fi
gzip -f $O/*_source.csv
rm -f $O/*.ncu-rep
du -sm $O
