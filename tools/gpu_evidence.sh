#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests, the default bench line, the ncu launch list of the
# same bench command and `ncu --set full` captures of the row-block kernels.  usage: tools/gpu_evidence.sh TAG [nopytest]
# gpurun copies back at most 64 MiB: the captures are kept small (a full-set kernel record is ~2.7 MB, more with source).
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
if [ "$2" != "nopytest" ]; then
  python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
  tail -3 $OUT/${TAG}_pytest.log
fi
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-cv --no-exact > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launches rc=$?"
# the dominant kernel with source (a launch of the tensor form: iterations 0 and 1 of the reference's start run the FP32 form)
ncu --set full --clock-control none --import-source on -k regex:'repulse_tc2_kernel' -s 4 -c 1 -f -o $OUT/prof_${TAG}_tc2 \
    python tools/gpu_rowblock_quick.py 100000 16 0.99 3 > $OUT/${TAG}_ncu_tc2.log 2>&1; echo "full tc2 rc=$?"
# the FP32 difference form as it runs in iteration 0
ncu --set full --clock-control none -k regex:'repulse_kernel' -s 0 -c 1 -f -o $OUT/prof_${TAG}_f32 \
    python tools/gpu_rowblock_quick.py 100000 16 0.99 3 > $OUT/${TAG}_ncu_f32.log 2>&1; echo "full f32 rc=$?"
# the other kernels of two iterations (one of them a check iteration)
ncu --set full --clock-control none -k regex:'spring_kernel|mae_kernel|combine_kernel|image_tc_kernel|image_t2_kernel' -s 12 -c 10 -f -o $OUT/prof_${TAG}_rest \
    python tools/gpu_rowblock_quick.py 100000 16 0.99 3 > $OUT/${TAG}_ncu_rest.log 2>&1; echo "full rest rc=$?"
du -sm $OUT
head -c 1500 $OUT/${TAG}_bench.json
