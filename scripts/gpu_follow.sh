cd $GRAFT_REPO_ROOT
for f in 1 2 4 8 64; do for c in rle8_3symlut rle8_7symlut; do echo "follow=$f"; HSRLE_FOLLOW=$f timeout 120 python scripts/prof_one.py $c 2 enc 2>&1 | tail -1 | cut -c1-300; done; done
