#!/bin/bash
# Same-box A/B of the epilogue refactor: libfemasr_old.so (previous commit) against the current library, with the
# compiled-in epilogue modes switched on selectively (FEMASR_EPI_MODES bit mask).  Outputs under gpurun_out/ab_*.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OLD=$PWD/femasr_b200/libfemasr_old.so
FEMASR_LIB=$OLD python scripts/ab_digest.py > gpurun_out/ab_digest_old.json 2> gpurun_out/ab_digest_old.err
python scripts/ab_digest.py > gpurun_out/ab_digest_new.json 2> gpurun_out/ab_digest_new.err
FEMASR_EPI_MODES=0 python scripts/ab_digest.py > gpurun_out/ab_digest_new_generic.json 2> gpurun_out/ab_digest_new_generic.err
cmp -s gpurun_out/ab_digest_old.json gpurun_out/ab_digest_new.json && echo DIGEST_EQUAL_new || echo DIGEST_DIFFER_new
cmp -s gpurun_out/ab_digest_old.json gpurun_out/ab_digest_new_generic.json && echo DIGEST_EQUAL_generic || echo DIGEST_DIFFER_generic
FEMASR_LIB=$OLD python scripts/profile_layers.py 32 > gpurun_out/ab_layers_old.txt 2>&1
for m in 15 0 1 3 12; do FEMASR_EPI_MODES=$m python scripts/profile_layers.py 32 > gpurun_out/ab_layers_new_m$m.txt 2>&1; done
FEMASR_LIB=$OLD python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_old.json 2> gpurun_out/ab_bench_old.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_new.json 2> gpurun_out/ab_bench_new.err
FEMASR_LIB=$OLD python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_old2.json 2> gpurun_out/ab_bench_old2.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/ab_bench_new2.json 2> gpurun_out/ab_bench_new2.err
for f in old new old2 new2; do python -c "import json;d=json.load(open('gpurun_out/ab_bench_$f.json'));print('$f',d['value'],d['ms_per_step'],d['clocks']['sm_mhz'])"; done
head -3 gpurun_out/ab_layers_old.txt gpurun_out/ab_layers_new_m*.txt | grep total
