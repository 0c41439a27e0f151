#!/bin/bash
# SASS evidence (no GPU needed): TMA bulk copies + mbarrier in predict_tiles_kernel, the shared-atomic inner loop and the
# cp.async ring of hist_stream_kernel.  Writes profiles/r02_sass_excerpt.md.
LIB=gbrl_b200/lib/libgbrl_b200.so
OUT=profiles/r02_sass_excerpt.md
TMP=$(mktemp)
cuobjdump -sass $LIB > $TMP
count() { awk -v k="$1" '/Function :/ {f=$0} index(f,k) {print}' $TMP | grep -c "$2"; }
{
echo "# SASS excerpts of libgbrl_b200.so (sm_100a), \`tools/sass_excerpt.sh\`"
echo
echo "Instruction counts per kernel (cuobjdump -sass):"
echo
echo "| kernel | mnemonic | count |"
echo "|---|---|---:|"
for k in "predict_tiles_kernelILi2" "predict_tiles_kernelILi4"; do
  for mn in "UBLKCP" "SYNCS.ARRIVE.TRANS64" "SYNCS.PHASECHK"; do echo "| $k | $mn | $(count $k $mn) |"; done
done
for k in "hist_stream_kernelILi1ELi24ELi3" "hist_stream_kernelILi2ELi24ELi2"; do
  for mn in "ATOMS" "LDGSTS" "LDGDEPBAR" "REDG\|RED.E" "STG"; do echo "| $k | $mn | $(count $k "$mn") |"; done
done
echo
echo "## predict_tiles_kernel<2>: TMA bulk copy of a tree chunk (cp.async.bulk.shared::cluster.global.mbarrier)"
echo '```'
awk '/Function :/ {f=$0} index(f,"predict_tiles_kernelILi2") {print}' $TMP | grep -B3 -A3 "UBLKCP" | head -40 | sed 's/^ *//' | cut -c1-150
echo '```'
echo
echo "## hist_stream_kernel<1,24,3>: shared-memory atomics of one (row, 32-feature tile) step"
echo '```'
awk '/Function :/ {f=$0} index(f,"hist_stream_kernelILi1ELi24ELi3") {print}' $TMP | grep -B2 -A2 "ATOMS" | head -60 | sed 's/^ *//' | cut -c1-150
echo '```'
} > $OUT
rm -f $TMP
wc -l $OUT
