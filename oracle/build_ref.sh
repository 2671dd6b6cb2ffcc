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

# ---- second oracle: the reference's OWN CUDA back-end (src/cuda/*.cu + cublasGemmEx), compiled unmodified for sm_100.
#   oracle/_ref/cuda/CIANNA.so   parity only (FP16 / BF16 behaviour of upstream itself, LRN), never timed, never shipped.
# Runs only on a GPU box (the tests that use it are -m gpu and skip when the file or libcublas is missing).
build_cuda_variant() {
	NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
	[ -x "$NVCC" ] || { echo "build_ref: no nvcc, skipping the cuda variant"; return 0; }
	CUDALIB=/usr/local/cuda/lib64
	OUT=$HERE/_ref/cuda; mkdir -p $OUT/obj
	CDEF="-D MAX_LAYERS_NB=200 -D MAX_NETWORKS_NB=10 -D CUDA_THREADS_PER_BLOCKS=256 -D CUDA"
	for s in cuda_main cuda_conv_layer cuda_dense_layer cuda_pool_layer cuda_norm_layer cuda_lrn_layer cuda_activ_functions; do
		$NVCC -O3 -w -arch=sm_100 -D GEN_AMPERE -D comp_CUDA $CDEF -Xcompiler -fPIC -I $REF -c $REF/cuda/$s.cu -o $OUT/obj/$s.o &
	done
	HSRCS="conv_layer.c dense_layer.c pool_layer.c norm_layer.c lrn_layer.c initializers.c vars.c auxil.c activ_functions.c
	       naiv/naiv_conv_layer.c naiv/naiv_dense_layer.c naiv/naiv_pool_layer.c naiv/naiv_norm_layer.c"
	for s in $HSRCS; do
		gcc -O3 -fPIC -std=c99 -w $CDEF -I $REF -c $REF/$s -o $OUT/obj/$(basename $s .c).o &
	done
	gcc -O3 -fPIC -std=c99 -w $CDEF -I $REF -I $PYINC -I $NPINC -c $REF/python_module.c -o $OUT/obj/python_module.o &
	gcc -O3 -fPIC -std=c99 -w $CDEF -I $REF -c $HERE/ref_probe_cuda.c -o $OUT/obj/ref_probe_cuda.o &
	wait
	for s in cuda_main cuda_conv_layer cuda_dense_layer cuda_pool_layer cuda_norm_layer cuda_lrn_layer cuda_activ_functions \
	         auxil activ_functions python_module ref_probe_cuda; do
		[ -f $OUT/obj/$s.o ] || { echo "build_ref: $s failed to compile (cuda variant)"; exit 1; }
	done
	g++ -shared -o $OUT/CIANNA.so $OUT/obj/*.o -lm -L $CUDALIB -lcublas -lcudart -lcurand -Wl,-rpath,$CUDALIB
	rm -rf $OUT/obj
	echo "built $OUT/CIANNA.so"
}
if [ "${CIANNA_REF_NO_CUDA:-0}" != "1" ]; then build_cuda_variant; fi

# ---- the product as upstream's back-end: upstream's UNMODIFIED host sources (-D CUDA) + cianna_b200/shim/cuda_b200_shim.c
# in place of src/cuda/*.cu, cuBLAS and cuRAND.   oracle/_ref/dropin/CIANNA.so   (tests/test_gpu_backends.py)
# Built under oracle/_ref/ because it contains compiled reference host code; it loads ../../../cianna_b200/libcianna_host.so.
build_dropin_variant() {
	REPO=$(cd "$HERE/.." && pwd)
	[ -f $REPO/cianna_b200/libcianna_host.so ] || { echo "build_ref: libcianna_host.so not built yet, skipping the dropin variant"; return 0; }
	OUT=$HERE/_ref/dropin; mkdir -p $OUT/obj
	CDEF="-D MAX_LAYERS_NB=200 -D MAX_NETWORKS_NB=10 -D CUDA_THREADS_PER_BLOCKS=256 -D CUDA"
	HSRCS="conv_layer.c dense_layer.c pool_layer.c norm_layer.c lrn_layer.c initializers.c vars.c auxil.c activ_functions.c
	       naiv/naiv_conv_layer.c naiv/naiv_dense_layer.c naiv/naiv_pool_layer.c naiv/naiv_norm_layer.c"
	for s in $HSRCS; do
		gcc -O3 -fPIC -std=c99 -w $CDEF -I $REF -c $REF/$s -o $OUT/obj/$(basename $s .c).o &
	done
	gcc -O3 -fPIC -std=c99 -w $CDEF -I $REF -I $PYINC -I $NPINC -c $REF/python_module.c -o $OUT/obj/python_module.o &
	gcc -O3 -fPIC -std=c99 -w $CDEF -I $REF -c $HERE/ref_probe_cuda.c -o $OUT/obj/ref_probe_cuda.o &
	gcc -O2 -fPIC -std=c99 -Wall -Wno-unused-function $CDEF -I $REF -I $REPO/include -I $REPO/cianna_b200/host \
		-c $REPO/cianna_b200/shim/cuda_b200_shim.c -o $OUT/obj/cuda_b200_shim.o &
	wait
	for s in auxil activ_functions python_module ref_probe_cuda cuda_b200_shim conv_layer dense_layer; do
		[ -f $OUT/obj/$s.o ] || { echo "build_ref: $s failed to compile (dropin variant)"; exit 1; }
	done
	gcc -shared -o $OUT/CIANNA.so $OUT/obj/*.o -lm -L $REPO/cianna_b200 -lcianna_host -lcianna_b200 \
		-Wl,-rpath,'$ORIGIN/../../../cianna_b200' -Wl,-Bsymbolic -Wl,--no-undefined $(python3-config --ldflags --embed 2>/dev/null || true)
	rm -rf $OUT/obj
	echo "built $OUT/CIANNA.so"
}
build_dropin_variant
