for t in 0 592 1184 2368 4736 9472; do
  echo "== SPLIT_TILES $t"
  C2M_WARP_SPLIT_TILES=$t python tools/prof_pyramid.py --graph-only --iters 30 2>&1 | grep -v done | cut -c1-90
done
C2M_WARP_SPLIT_TILES=9472 python tools/stress.py 150 5
