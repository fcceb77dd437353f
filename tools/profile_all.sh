#!/bin/bash
# One gpurun call: plain runs first (must exit 0), then ncu captures, every ncu under its own timeout and
# restricted to the dae kernels (ncu cannot replay some library GEMMs of the PyTorch encoder and then hangs).
# Usage: tools/profile_all.sh <tag> [families...]     Outputs land in gpurun_out/.
set -u
TAG=${1:-r01}; shift
FAMS=${@:-ctc greedy specaug stitch softdtw beam}
O=gpurun_out
mkdir -p $O
KRE='regex:ctc_|argmax_rows|collapse_kernel|greedy_fused|specaug_|window_sums|stitch_kernel|softdtw_|beam_search|cutout_|ngram_|frame_shuffle|noise_'
for fam in $FAMS; do
  extra=""; [ "$fam" = ctc ] && extra="--with-scale"      # the adapt step's call: dae_ctc_loss_grad
  timeout 300 python tools/prof_one.py $fam --reps 2 $extra > $O/plain_$fam.log 2>&1 || { echo "plain run of $fam failed"; tail -5 $O/plain_$fam.log; exit 1; }
done
timeout 300 python bench.py --steps 1 --warmup 1 --frames 40000 --no-aux > $O/plain_bench.log 2>&1 || { echo "plain bench failed"; tail -5 $O/plain_bench.log; exit 1; }
# launch list of the bench command, dae kernels only: device time of every launch (cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KRE" -c 400 --csv \
    --log-file $O/launches_${TAG}.csv python bench.py --steps 1 --warmup 1 --frames 40000 --no-aux > $O/ncu_bench.log 2>&1
echo "launch list rc=$?"
for fam in $FAMS; do
  extra=""; [ "$fam" = ctc ] && extra="--with-scale"
  timeout 420 ncu --set full --clock-control none --import-source on -k "$KRE" -c 8 -f -o $O/${fam}_${TAG} \
      python tools/prof_one.py $fam --reps 1 $extra > $O/ncu_$fam.log 2>&1
  echo "$fam rc=$?"
done
ls -la $O | grep -E "ncu-rep|launches"
