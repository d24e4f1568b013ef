#!/bin/sh
# ORACLE — builds the UNMODIFIED reference (every .c named by /root/reference/Makefile:27-83)
# straight from where it lies, links it with oracle/kmc_cpu.c in place of the absent
# libs/KMC/libkmc.a, and leaves only binaries under oracle/_ref/.  Nothing is copied from
# the reference tree.  The reference's own build system is NOT run (its Makefile needs git
# and libkmc.a); this is the "gcc on those files directly" recipe.
#
#   TA_ref      reference CPU path: ./TA_ref build_0 -1 R1.fq -2 R2.fq -l ust -k0 31 -t 8 -o out
#   TA_gpu      same objects, but build_initial_graph / build_graph_from_scratch(_without_count)
#               and KMC_build_kmer_database resolve to libtagpu.so (drop-in test; only built when
#               ../turingassembler_b200/libtagpu.so exists)
#   TA_kmc      all reference objects unmodified, libtagpu.so only replaces libkmc.a
#               (KMC_build_kmer_database on the GPU, the reference's own reader and graph code after it)
set -e
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
if [ ! -d "$REF/src" ]; then
	echo "build_ref.sh: $REF not present; keeping any prebuilt $OUT" >&2
	exit 0
fi
mkdir -p "$OUT/obj"
SRCS=$(sed -n '/^SRC *=/,/^OBJ *=/p' "$REF/Makefile" | grep -o 'src/[A-Za-z0-9_/]*\.c' | sort -u)
CFLAGS="-std=gnu99 -m64 -O2 -w -DLOG_USE_COLOR -DGIT_SHA=\"oracle\" -I $REF -I $REF/src -I $REF/src/minimizers -fPIC -g -pthread"
NPROC=$(nproc 2>/dev/null || echo 4)
pids=""
n=0
for s in $SRCS; do
	o=$OUT/obj/$(echo "$s" | tr '/' '_' | sed 's/\.c$/.o/')
	if [ ! -f "$o" ] || [ "$REF/$s" -nt "$o" ]; then
		gcc $CFLAGS -c "$REF/$s" -o "$o" &
		pids="$pids $!"
		n=$((n + 1))
		if [ $((n % NPROC)) -eq 0 ]; then wait $pids; pids=""; fi
	fi
done
wait $pids
gcc -std=gnu99 -O2 -g -fPIC -pthread -I "$HERE" -c "$HERE/kmc_cpu.c" -o "$OUT/obj/kmc_cpu_shim.o"
LIBS="$REF/libs/zlib/libz.a $REF/libs/bzip2/libbz2.a $REF/libs/bwa/libbwa.a -lm"
g++ -pthread -o "$OUT/TA_ref" $(ls "$OUT"/obj/src_*.o) "$OUT/obj/kmc_cpu_shim.o" $LIBS
echo "built $OUT/TA_ref"
# build_local_assembly_graph has no sub-command of its own: our 40-line driver + the unmodified reference objects
gcc $CFLAGS -c "$HERE/local_ref_main.c" -o "$OUT/obj/local_ref_main.o"
g++ -pthread -o "$OUT/TA_local_ref" $(ls "$OUT"/obj/src_*.o | grep -v src_main.o) "$OUT/obj/local_ref_main.o" "$OUT/obj/kmc_cpu_shim.o" $LIBS
echo "built $OUT/TA_local_ref"
# the contig-file mode of build_graph_from_scratch (n_files < 0) is set by no sub-command either: same recipe
gcc $CFLAGS -c "$HERE/contig_ref_main.c" -o "$OUT/obj/contig_ref_main.o"
g++ -pthread -o "$OUT/TA_contig_ref" $(ls "$OUT"/obj/src_*.o | grep -v src_main.o) "$OUT/obj/contig_ref_main.o" "$OUT/obj/kmc_cpu_shim.o" $LIBS
echo "built $OUT/TA_contig_ref"

TAGPU=$HERE/../turingassembler_b200/libtagpu.so
if [ -f "$TAGPU" ]; then
	# Drop-in link: hide the three stage entry points inside the reference's kmer_build.o so every
	# other reference object binds them (and KMC_build_kmer_database) to libtagpu.so instead.
	objcopy --localize-symbol=build_initial_graph \
		--localize-symbol=build_graph_from_scratch \
		--localize-symbol=build_graph_from_scratch_without_count \
		"$OUT/obj/src_kmer_build.o" "$OUT/obj/dropin_kmer_build.o"
	OBJS=$(ls "$OUT"/obj/src_*.o | grep -v src_kmer_build.o)
	g++ -pthread -o "$OUT/TA_gpu" $OBJS "$OUT/obj/dropin_kmer_build.o" \
		-L "$HERE/../turingassembler_b200" -ltagpu -Wl,-rpath,'$ORIGIN/../../turingassembler_b200' $LIBS
	echo "built $OUT/TA_gpu"
	# Library-boundary drop-in (INTEGRATION.md option B): ALL reference objects untouched — its own
	# build_graph_from_scratch, KMC_reader, kmhash ... — and libtagpu.so only in place of libkmc.a, i.e. the GPU
	# writes the KMC database and the reference's reader / graph builder consume it.
	g++ -pthread -o "$OUT/TA_kmc" $(ls "$OUT"/obj/src_*.o) \
		-L "$HERE/../turingassembler_b200" -ltagpu -Wl,-rpath,'$ORIGIN/../../turingassembler_b200' $LIBS
	echo "built $OUT/TA_kmc"
	# drop-in for the local-assembly re-entry
	objcopy --localize-symbol=build_local_assembly_graph "$OUT/obj/dropin_kmer_build.o" "$OUT/obj/dropin_local_kmer_build.o"
	g++ -pthread -o "$OUT/TA_local_gpu" $(echo "$OBJS" | tr ' ' '\n' | grep -v src_main.o) "$OUT/obj/dropin_local_kmer_build.o" \
		"$OUT/obj/local_ref_main.o" -L "$HERE/../turingassembler_b200" -ltagpu -Wl,-rpath,'$ORIGIN/../../turingassembler_b200' $LIBS
	echo "built $OUT/TA_local_gpu"
	g++ -pthread -o "$OUT/TA_contig_gpu" $(echo "$OBJS" | tr ' ' '\n' | grep -v src_main.o) "$OUT/obj/dropin_kmer_build.o" \
		"$OUT/obj/contig_ref_main.o" -L "$HERE/../turingassembler_b200" -ltagpu -Wl,-rpath,'$ORIGIN/../../turingassembler_b200' $LIBS
	echo "built $OUT/TA_contig_gpu"
fi
