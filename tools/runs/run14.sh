cd $GRAFT_REPO_ROOT
WL=1080 bash tools/variants.sh "-DXYB_PAIR_INLINE=__noinline__" "-DXYB_PAIR_INLINE=__forceinline__" 2>&1 | tail -8
