#!/bin/bash
# profiles/make_sass_listing.sh : cuobjdump -sass of the headline instantiation of the streaming loss kernel
# (focal, fused select, float32, compact labels) + opcode histogram -> profiles/r02_sass_cls_tile.txt
set -e
ROOT=$(cd $(dirname $0)/.. && pwd)
OBJ=$ROOT/retinanet_tensorflow_b200/lib/rn_loss_tile.o
FUN='_ZN2rn15cls_tile_kernelILi1ELb0ELb1ELb0ELb0EEEvNS_4SegsENS_7TileDevE'
OUT=$ROOT/profiles/r02_sass_cls_tile.txt
{
  echo "# cuobjdump -sass -fun cls_tile_kernel<RN_CLS_FOCAL, dice=0, select=1, bf16=0, dense=0> (sm_100a)"
  echo "# built by: make -C retinanet_tensorflow_b200/csrc ; listing by: profiles/make_sass_listing.sh"
  grep -A3 "Compiling entry function '$FUN'" $ROOT/retinanet_tensorflow_b200/lib/rn_loss_tile.ptxas.log | sed 's/^/# /'
  echo "#"
  echo "# opcode histogram (mnemonic before the first '.', count):"
  cuobjdump -sass -fun "$FUN" $OBJ | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's#^\s+/\*[0-9a-f]+\*/\s+##' \
    | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/\..*//; s/;//' | sort | uniq -c | sort -rn \
    | awk '{printf "#   %-12s %d\n", $2, $1}'
  echo "#"
  echo "# Blackwell-specific evidence: UBLKCP (cp.async.bulk global->shared), SYNCS (mbarrier), FFMA2/FMUL2/FADD2 (packed f32x2)"
  echo "#"
  cuobjdump -sass -fun "$FUN" $OBJ | grep -E '^\s+/\*[0-9a-f]{4}\*/' | sed -E 's#\s*/\* 0x[0-9a-f]+ \*/\s*$##'
} > $OUT
wc -l $OUT
