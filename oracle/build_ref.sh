#!/bin/bash
# build_ref.sh -- TEST INFRASTRUCTURE ONLY (oracle).
#
# Compiles the UNMODIFIED reference (SLAM++) from the sources where they lie under
# $SPP_REFERENCE (default /root/reference) into oracle/_ref/, together with the thin drivers
# oracle/ref_driver_*.cpp. Nothing is copied into the repository; oracle/_ref/ is git-ignored
# (but travels to the GPU box with gpurun so the CPU baseline can be timed there).
#
# The reference's own CMake build is not used (CMake >= 4 rejects CMakeLists.txt:5, policy
# CMP0014 OLD). This is the direct-g++ recipe of SURVEY.md F2: one flag set for all TUs
# (mixing -march settings breaks Eigen's alignment ABI), /usr/bin/g++ (the image's $CC gcc
# has no libgomp spec).
#
# usage: oracle/build_ref.sh [driver ...]      (default drivers: ba dropin pose order dropin_pose parse dropin_lm dropin_gn)
set -u
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${SPP_REFERENCE:-/root/reference}"
OUT="${SPP_REF_OUT:-$HERE/_ref}"
OBJ="$OUT/obj"
CXX=/usr/bin/g++
CC=/usr/bin/gcc
JOBS="${SPP_JOBS:-8}"
# x86-64-v3 = AVX2+FMA: runs on this container and on any B200 host CPU
OPT="${SPP_REF_OPT:--O3 -march=x86-64-v3}"
DEF="-DNDEBUG -D_UNIX -D__DISABLE_GPU -DEIGEN_DONT_PARALLELIZE -DNTIMER -DNPARTITION -DDLONG"
EIGEN="${SPP_REF_EIGEN:-32}" # the reference's CMake default is Eigen 3.2 (CMakeLists.txt:179, SLAM_P_P_EIGEN33 FALSE)
INC="-I$REF/include -I$REF/include/eigen$EIGEN -I$REF/include/cholmod -I$REF/include/cholmod/AMD \
 -I$REF/include/cholmod/CAMD -I$REF/include/cholmod/CCOLAMD -I$REF/include/cholmod/COLAMD \
 -I$REF/include/cholmod/SuiteSparse -I$HERE"

if [ ! -d "$REF/include/slam" ]; then
	echo "build_ref: reference not present at $REF (fine on the GPU box: prebuilt oracle/_ref is used)"
	exit 0
fi
mkdir -p "$OBJ"
DRIVERS=("$@")
[ ${#DRIVERS[@]} -eq 0 ] && DRIVERS=(ba dropin pose order dropin_pose parse dropin_lm dropin_gn)

compile_one() { # src obj compiler extra
	local src="$1" obj="$2" comp="$3"; shift 3
	if [ -f "$obj" ] && [ "$obj" -nt "$src" ]; then return 0; fi
	$comp $OPT -fopenmp $DEF $INC "$@" -c "$src" -o "$obj" 2> "$obj.log" || { echo "FAIL $src (see $obj.log)"; return 1; }
}
export -f compile_one
export OPT DEF INC CXX CC

LIST="$OBJ/jobs.txt"
: > "$LIST"
for f in BlockMatrix Debug LinearSolver_Schur LinearSolver_Schur_GPU LinearSolver_CSparse OrderingMagic Parser Tags Tga Timer; do
	echo "$REF/src/slam/$f.cpp $OBJ/slam_$f.o $CXX" >> "$LIST"
done
for d in "${DRIVERS[@]}"; do
	# the drop-in driver also sees the product's public headers (C ABI + reference-side adapter): a driver object older
	# than any of them (or than the dump helper) is stale
	if [ -f "$OBJ/ref_driver_$d.o" ] && [ -n "$(find "$HERE/../include" "$HERE/spp_dump.h" -newer "$OBJ/ref_driver_$d.o" -type f 2>/dev/null | head -1)" ]; then
		rm -f "$OBJ/ref_driver_$d.o"
	fi
	echo "$HERE/ref_driver_$d.cpp $OBJ/ref_driver_$d.o $CXX -I$HERE/../include" >> "$LIST"
done
for f in "$REF"/src/csparse/*.c; do
	echo "$f $OBJ/csparse_$(basename "$f" .c).o $CC -I$REF/include/csparse" >> "$LIST"
done
for f in "$REF"/src/cholmod/AMD/*.c; do
	echo "$f $OBJ/amd_$(basename "$f" .c).o $CC" >> "$LIST"
done
for f in "$REF"/src/cholmod/CAMD/*.c; do
	echo "$f $OBJ/camd_$(basename "$f" .c).o $CC" >> "$LIST"
done
# heavy C++ TUs first so they overlap with the many small C files
xargs -P "$JOBS" -L 1 bash -c 'compile_one "$@"' _ < "$LIST" || { echo "build_ref: compilation failed"; exit 1; }

LIBOBJ=$(ls "$OBJ"/slam_*.o "$OBJ"/csparse_*.o "$OBJ"/amd_*.o "$OBJ"/camd_*.o)
rc=0
for d in "${DRIVERS[@]}"; do
	EXTRA=""
	if [ "$d" = "dropin" ] || [ "$d" = "dropin_pose" ] || [ "$d" = "dropin_lm" ] || [ "$d" = "dropin_gn" ]; then # links the product library; found at run time relative to the binary
		EXTRA="-L$HERE/../slam_plus_plus_b200 -lspp_b200 -Wl,-rpath,\$ORIGIN/../../slam_plus_plus_b200"
	fi
	$CXX -fopenmp -o "$OUT/ref_driver_$d" "$OBJ/ref_driver_$d.o" $LIBOBJ -lrt $EXTRA || rc=1
done
[ $rc -eq 0 ] && echo "build_ref: ok -> $OUT" || echo "build_ref: link failed"
exit $rc
