#!/bin/bash
# Same-box A/B of the merged tap-row MMAs in the resident-weight 64 -> 64 strip kernel (FEMASR_BRES_MERGE).
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
OLD=$PWD/femasr_b200/libfemasr_nomerge.so
FEMASR_LIB=$OLD python scripts/ab_digest.py > gpurun_out/abm_digest_old.json 2> gpurun_out/abm_digest_old.err
timeout 300 python scripts/ab_digest.py > gpurun_out/abm_digest_new.json 2> gpurun_out/abm_digest_new.err
FEMASR_F8_CROSS=0 FEMASR_LIB=$OLD python scripts/ab_digest.py > gpurun_out/abm_digest_old_nof8.json 2>> gpurun_out/abm_digest_old.err
FEMASR_F8_CROSS=0 python scripts/ab_digest.py > gpurun_out/abm_digest_new_nof8.json 2>> gpurun_out/abm_digest_new.err
cmp -s gpurun_out/abm_digest_old.json gpurun_out/abm_digest_new.json && echo DIGEST_EQUAL_f8 || echo DIGEST_DIFFER_f8
cmp -s gpurun_out/abm_digest_old_nof8.json gpurun_out/abm_digest_new_nof8.json && echo DIGEST_EQUAL_nof8 || echo DIGEST_DIFFER_nof8
timeout 300 python -m pytest tests/test_tc_gpu.py -m gpu -q -x 2>&1 | tail -3
for t in old new; do
  L=""; [ $t = old ] && L=$OLD
  FEMASR_LIB=$L MB_ONLY="64->64" MB_F8=1 MB_GN=1 python scripts/microbench_tc.py > gpurun_out/abm_mb_f8_$t.txt 2>&1
  FEMASR_LIB=$L MB_ONLY="64->64" MB_GN=1 python scripts/microbench_tc.py > gpurun_out/abm_mb_$t.txt 2>&1
done
FEMASR_LIB=$OLD python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abm_bench_old.json 2> gpurun_out/abm_bench_old.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abm_bench_new.json 2> gpurun_out/abm_bench_new.err
FEMASR_LIB=$OLD python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abm_bench_old2.json 2> gpurun_out/abm_bench_old2.err
python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/abm_bench_new2.json 2> gpurun_out/abm_bench_new2.err
grep -h "64->64" gpurun_out/abm_mb_f8_old.txt gpurun_out/abm_mb_f8_new.txt gpurun_out/abm_mb_old.txt gpurun_out/abm_mb_new.txt
for f in old new old2 new2; do python -c "import json;d=json.load(open('gpurun_out/abm_bench_$f.json'));print('$f',d['value'],d['ms_per_step'],d['clocks']['sm_mhz'])"; done
