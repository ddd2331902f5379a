#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests, the default bench line, the ncu launch list of the
# same bench command and one `ncu --set full` capture of the row-block kernels.  usage: tools/gpu_evidence.sh TAG
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/${TAG}_pytest.log
tail -3 $OUT/${TAG}_pytest.log
python bench.py > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e --no-cv --no-exact > $OUT/${TAG}_ncu_launches.log 2>&1; echo "launches rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'repulse_tc2_kernel|repulse_kernel|image_tc_kernel|image_t2_kernel|spring_kernel|mae_kernel|combine_kernel' \
    -s 18 -c 18 -f -o $OUT/prof_${TAG} python tools/gpu_rowblock_quick.py 100000 16 0.99 3 > $OUT/${TAG}_ncu_full.log 2>&1; echo "full rc=$?"
head -c 1500 $OUT/${TAG}_bench.json
