#!/bin/bash
# ncu evidence of the current tree (run under gpurun; everything lands in gpurun_out/).
#   tools/capture_profiles.sh <tag>      e.g. r02_v7
# 1. launch list of a short bench run (gpu__time_duration.sum)
# 2. DRAM bytes of every launch of exactly one warm full-size search (tools/one_search.py) -> the `traffic` fields of the bench line
# 3. `--set full` captures of the lean tier (the two largest launches of a warm search), the dense seed scan
#    (k_seed_scan_smem) and the region scan
tag=${1:-r02}
out=gpurun_out
mkdir -p $out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv --log-file $out/launches_$tag.csv \
	python bench.py --mbp 200 --steps 2 --warmup 1 --no-cpu-baseline > $out/launches_$tag.log 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --profile-from-start off \
	--csv --log-file $out/dram_$tag.csv python tools/one_search.py 1000 100 taqman > $out/dram_$tag.log 2>&1
$NCU --set full --import-source on -k regex:'k_align_fast' -s 36 -c 2 -f -o $out/ncu_k_align_lean_$tag \
	python bench.py --mbp 200 --steps 1 --warmup 1 --no-cpu-baseline --no-fasta > $out/ncu_lean_$tag.log 2>&1
$NCU --set full --import-source on -k k_seed_scan_smem -s 2 -c 1 -f -o $out/ncu_k_seed_scan_smem_$tag \
	python bench.py --mbp 200 --steps 1 --warmup 1 --no-cpu-baseline --no-fasta > $out/ncu_scan_$tag.log 2>&1
$NCU --set full --import-source on -k k_region_scan -s 3 -c 1 -f -o $out/ncu_k_region_scan_$tag \
	python bench.py --mbp 200 --steps 1 --warmup 1 --no-cpu-baseline --no-fasta > $out/ncu_region_$tag.log 2>&1
ls -la $out/*$tag*
