#!/bin/bash
# tuning helper: build liblsf with extra compile flags into variants/NAME.so (separate object dir).  usage: tools/build_variant.sh NAME "-DFLAG ..."
set -e
NAME=$1; EXTRA=$2
ROOT=$(cd $(dirname $0)/.. && pwd)
B=/tmp/lsf_variant_$NAME
rm -rf $B && mkdir -p $B/levelsetfortran_b200 $B/include $B/tests $ROOT/variants
cp -r $ROOT/levelsetfortran_b200/csrc $B/levelsetfortran_b200/ && cp $ROOT/include/*.h $B/include/ && cp -r $ROOT/tests/emu $B/tests/
rm -f $B/levelsetfortran_b200/csrc/*.o
make -C $B/levelsetfortran_b200/csrc -j8 EXTRA="$EXTRA" TARGET=$ROOT/variants/$NAME.so > $B/make.log 2>&1 || (tail -20 $B/make.log; false)
echo built variants/$NAME.so
