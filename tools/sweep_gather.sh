for h in 0 1 2 3 4 5 7; do CBN_GATHER_TILES=1 CBN_GT_HINTS=$h python tools/exp_gather2.py; done
for st in 2 3 6 8; do CBN_GATHER_TILES=1 CBN_GT_STAGES=$st python tools/exp_gather2.py; done
for ps in 1 2 3; do CBN_GATHER_TILES=1 CBN_GT_PER_SM=$ps python tools/exp_gather2.py; done
CBN_GATHER_TILES=0 python tools/exp_gather2.py
