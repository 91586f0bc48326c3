cd $GRAFT_REPO_ROOT
bash tools/sweep.sh w8 w8s3 w4 w6t15s5 w16b1t40 w8m2 w8s2b3 w8b1s8
