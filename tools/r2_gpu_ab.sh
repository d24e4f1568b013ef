#!/bin/sh
# developer helper (runs ON the GPU box): [parity tests on the product library,] then the per-kernel breakdown of one build per
# workload with each of the given library variants (turingassembler_b200/libtagpu<suffix>.so), same box, interleaved
# usage: r2_gpu_ab.sh <tag> <test: 0|1> "<workloads>" <suffix> ...   ("" = product library)
tag=${1:-r2x}; dotest=${2:-1}; wls=${3:-C2}; shift 3
if [ "$dotest" = 1 ]; then
	python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1
	tail -5 gpurun_out/${tag}_pytest.log
fi
for rep in 1 2; do
	for suf in "$@"; do
		[ "$suf" = "-" ] && suf=""
		lib=libtagpu${suf}.so
		[ -f turingassembler_b200/$lib ] || continue
		for wl in $wls; do
			TAGPU_LIB=$PWD/turingassembler_b200/$lib python tools/prof_run.py $wl 3 > gpurun_out/${tag}_${lib}_${wl}_${rep}.log 2>&1
			echo "== $lib $wl rep $rep"; grep -A4 "^count " gpurun_out/${tag}_${lib}_${wl}_${rep}.log | head -5
		done
	done
done
if [ -f turingassembler_b200/libtagpu_timing.so ]; then
	TAGPU_LIB=$PWD/turingassembler_b200/libtagpu_timing.so python tools/prof_run.py C2 1 2>&1 | grep "tagpu timing" | tail -1
fi
