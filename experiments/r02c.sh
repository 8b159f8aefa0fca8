#!/bin/bash
for knobs in "" "HPVG_FUSED_BN=0" "HPVG_FUSE_MASK=0" "HPVG_CRITIC_WSIDE=0" "HPVG_WGRAD_STACK=0" "HPVG_FUSED_BN=0 HPVG_FUSE_MASK=0 HPVG_CRITIC_WSIDE=0 HPVG_WGRAD_STACK=0"; do
  echo "== knobs: [$knobs]"
  env $knobs timeout 100 python experiments/sg_wide_knobs.py train_sg_wide 2>&1 | tail -5
done
echo "== gan wide"; timeout 100 python experiments/sg_wide_knobs.py train_gan_wide 2>&1 | tail -5
