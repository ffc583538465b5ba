#!/bin/bash
# One GPU-box visit: tests, bench lines, launch list, ncu --set full of the step's kernels.
#   gpurun --timeout 1700 -- 'bash tools/gpu_round.sh TAG [tests] [bench] [ref] [launches] [full] [gz]'
# Every output goes to gpurun_out/TAG_*; tools/ncu_summary.py turns the reports into the text kept in profiles/.
set -u
tag="$1"; shift
what=" $* "
out=gpurun_out
mkdir -p $out
if [[ "$what" == *" tests "* ]]; then
    python -m pytest tests -m gpu -x -q --durations=8 > $out/${tag}_pytest.log 2>&1
    echo "pytest rc=$?"; tail -4 $out/${tag}_pytest.log
fi
if [[ "$what" == *" bench "* ]]; then
    python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench.err
    echo "bench rc=$?"; cut -c1-400 $out/${tag}_bench_1gpu.json
fi
if [[ "$what" == *" ref "* ]]; then
    python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2>> $out/${tag}_bench.err
    echo "ref rc=$?"; cut -c1-300 $out/${tag}_bench_ref.json
fi
small="--reads 262144 --steps 2 --warmup 1 --resident 1 --no-e2e --no-cpu-baseline --no-extra"
if [[ "$what" == *" launches "* ]]; then
    python bench.py $small > /dev/null 2>&1 && \
    ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
        python bench.py $small > $out/${tag}_ncu_launches.log 2>&1
    echo "launches rc=$?"
fi
if [[ "$what" == *" full "* ]]; then
    # one launch of every kernel of the step (the second step: skip the warm-up step's launches)
    timeout 900 ncu --set full --clock-control none --import-source on \
        -k regex:'seed_kernel|trigger_kernel|filter_kernel|scan_kernel|resolve_band_kernel|emit_kernel|pack_kernel|select_kernel|bucket_scatter' \
        --launch-skip 18 -c 18 -o $out/${tag}_full -f python bench.py $small > $out/${tag}_ncu_full.log 2>&1
    echo "full rc=$?"
fi
if [[ "$what" == *" gz "* ]]; then
    # the gzip stage (orc_params.emit_gzip): device times, launch list, one ncu --set full launch of its kernels
    for a in "262144" "262144 all" "1048576"; do python tools/gz_probe.py $a; done > $out/${tag}_gz_probe.json 2> $out/${tag}_gz.err
    echo "gz probe rc=$?"; cat $out/${tag}_gz_probe.json
    ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gz_|emit_kernel' --csv --log-file $out/${tag}_gz_launches.csv \
        python tools/gz_probe.py 262144 > /dev/null 2>> $out/${tag}_gz.err
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gz_' --launch-skip 18 -c 6 \
        -o $out/${tag}_gz_full -f python tools/gz_probe.py 262144 > $out/${tag}_ncu_gz_full.log 2>&1
    echo "gz full rc=$?"
fi
