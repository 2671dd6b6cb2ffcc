#!/bin/sh
# Builds the UNMODIFIED reference CPU back-ends (C_NAIV + C_BLAS) straight from
# /root/reference/src into oracle/_ref/ (git-ignored, ships to the GPU box as a .so).
# TEST INFRASTRUCTURE ONLY: the product never loads anything from oracle/.
#
#   oracle/_ref/omp/CIANNA.so     upstream flags (-D BLAS -D OPEN_MP), used as timed CPU baseline
#   oracle/_ref/serial/CIANNA.so  no OpenMP, activ_functions.c built with -Dabs=fabsf
#                                 (SURVEY.md 8c caveats vii/viii: CUDA float semantics for YOLO)
# Both also contain oracle/ref_probe.c (our per-mini-batch driver / tensor reader).
# No reference source is copied: every -c reads the file where it lies.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
REF=${CIANNA_REF_SRC:-/root/reference/src}
if [ ! -d "$REF" ]; then echo "build_ref: $REF not present, skipping"; exit 0; fi
SP=$(python3 -c 'import sysconfig;print(sysconfig.get_paths()["purelib"])')
OBDIR=$SP/opencv_python_headless.libs
OB=$(ls $OBDIR/libopenblasp-r0-*.so | head -1)
PYINC=$(python3 -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
NPINC=$(python3 -c 'import numpy;print(numpy.get_include())')
DEF="-D MAX_LAYERS_NB=200 -D MAX_NETWORKS_NB=10 -D BLAS"
OPT="-O3 -fPIC -std=c99 -w -I $HERE/shim -I $REF"
SRCS="conv_layer.c dense_layer.c pool_layer.c norm_layer.c lrn_layer.c initializers.c vars.c auxil.c
      naiv/naiv_conv_layer.c naiv/naiv_dense_layer.c naiv/naiv_pool_layer.c naiv/naiv_norm_layer.c
      blas/blas_conv_layer.c blas/blas_dense_layer.c"
build_variant() {
	V=$1; EXTRA=$2; ACTIV_EXTRA=$3
	OUT=$HERE/_ref/$V; mkdir -p $OUT/obj
	for s in $SRCS; do
		o=$OUT/obj/$(basename $s .c).o
		gcc $OPT $DEF $EXTRA -c $REF/$s -o $o &
	done
	gcc $OPT $DEF $EXTRA $ACTIV_EXTRA -c $REF/activ_functions.c -o $OUT/obj/activ_functions.o &
	gcc $OPT $DEF $EXTRA -I $PYINC -I $NPINC -c $REF/python_module.c -o $OUT/obj/python_module.o &
	gcc $OPT $DEF $EXTRA -c $HERE/ref_probe.c -o $OUT/obj/ref_probe.o &
	wait
	for s in $SRCS activ_functions.c python_module.c ref_probe.c; do
		[ -f $OUT/obj/$(basename $s .c).o ] || { echo "build_ref: $s failed to compile"; exit 1; }
	done
	gcc -shared -o $OUT/CIANNA.so $OUT/obj/*.o -lm $EXTRA $OB -Wl,-rpath,$OBDIR
	rm -rf $OUT/obj
	echo "built $OUT/CIANNA.so"
}
build_variant omp "-fopenmp -D OPEN_MP" ""
build_variant serial "" "-include $HERE/shim/abs_float.h"
