for mode in grouped classic; do
  export PPF_B200_VOTE=$mode
  echo "== $mode"
  python tools/profile_vote.py 10000 50000 8 2 2>&1 | tail -1
  python tools/profile_vote.py 1000 1000 1 3 2>&1 | tail -1
  python tools/profile_vote.py 2000 16000 1 2 2>&1 | tail -1
  python tools/profile_vote.py 5000 20000 4 2 2>&1 | tail -1
  python tools/profile_big_scene.py 1000000 8 2>&1 | tail -2
done
