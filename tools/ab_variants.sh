#!/bin/bash
# A/B of prebuilt library variants on the GPU box: every nanopore-barcoding-orc_b200/_build/variants/lib_*.so takes the
# place of the in-tree library for one short device-resident bench run (BENCH_ARGS to change it); the in-tree
# library is put back afterwards.  Output: one line per variant in gpurun_out/ab_<tag>.txt.
tag="${1:-ab}"
pkg=nanopore-barcoding-orc_b200
lib=$pkg/orcdemux/liborcdemux.so
args="${BENCH_ARGS:---steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-extra}"
mkdir -p gpurun_out
cp $lib /tmp/lib_keep.so
: > gpurun_out/ab_$tag.txt
for v in $pkg/_build/variants/lib_*.so; do
    cp "$v" $lib
    python bench.py $args 2> /tmp/ab.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['stages_ms_last_step']['kernel_ms']
print('$(basename $v)', round(d['value']/1e6,2), 'Mreads/s', round(d['ms_per_step'],3), 'ms', ' '.join('%s=%s'%(n,'+'.join('%.3f'%kk[n] for kk in k)) for n in k[0]), 'emit=%.3f'%d['stages_ms_last_step']['emit_ms'])
" >> gpurun_out/ab_$tag.txt 2>&1 || tail -3 /tmp/ab.err >> gpurun_out/ab_$tag.txt
done
cp /tmp/lib_keep.so $lib
cat gpurun_out/ab_$tag.txt
