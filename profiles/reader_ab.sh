#!/bin/bash
# update-iteration kernel times of the in-tree lib and of every lib in lib/variants (reader tile size experiments)
for so in 3drvo-marl-collisionavoidance_b200/lib/libuaisa_b200.so 3drvo-marl-collisionavoidance_b200/lib/variants/*.so; do
  echo "== $(basename $so)"; UAISA_B200_LIB=$PWD/$so python profiles/ppo_update_profile.py 2>/dev/null | grep -E "====|reader" | cut -c1-100
done
