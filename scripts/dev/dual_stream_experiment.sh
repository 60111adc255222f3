#!/bin/bash
# Record of the round-2 co-scheduling experiment (the code it drove was removed after the measurement, commit "dual-stream
# experiment"): halves of a launch group on two streams half an iteration apart (ASR_DUAL=1), optionally with K2's shared
# memory inflated by 28 KB so that only one K2 CTA fits an SM and K1 CTAs of the other half take the rest (ASR_K2_PAD_KB=28).
# bench.py --images 64, images/s and us per image-iteration (K1, K2 as seen by CUDA events on their own streams):
#   base                 61.83   K1 23.94  K2 29.80
#   dual                 62.61   (kernels overlap: per-kernel times double, throughput +1.3 %)
#   dual + pad 28 KB     52.26
#   single + pad 28 KB   51.41   K1 24.10  K2 40.56   (K2 with one CTA per SM is 36 % slower)
# Both kernels keep the issue slots 65-70 % busy on their own; there is little idle capacity for the other to use.
