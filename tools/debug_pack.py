import os, sys, random
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cse305_parallel_sequence_alignment_b200 as psa
from oracle import pyoracle as po
from tests.helpers import py_random_pair
ctx = psa.Context(0)
rnd = random.Random(3)
for (g,h,mode,alpha,maxm,maxn) in [(1,2,0,b"ACGT",40,48),(1,2,0,b"ACGT",96,128),(1,2,1,b"ACGT",96,128),(2,1,0,b"ACGT",150,150),(1,2,0,b"ACGTN",40,48),(0,2,0,b"ACGT",30,30),(1,0,1,b"ACGT",200,256)]:
    pairs=[py_random_pair(rnd,maxm,maxn,alpha) for _ in range(200)]
    pairs=[(a,b[:maxn]) for a,b in pairs]
    ba,oa,la = psa.pack_pairs([a for a,b in pairs]); bb,ob,lb = psa.pack_pairs([b for a,b in pairs])
    items, ops = ctx.align_batch(ba,oa,la,bb,ob,lb,mode,g,h,traceback=True)
    bad=0
    for k,(a,b) in enumerate(pairs):
        w = po.align(a,b,g,h,mode=mode)
        it = items[k]
        got_ops = psa.unpack_ops(ops[k], int(it["aln_len"]))
        ok = it["score"]==w.score and (mode==1 or (it["t1"],it["t2"],it["t3"],it["end_state"])==(w.t1,w.t2,w.t3,w.end_state)) and got_ops==w.ops and (it["end_i"],it["end_j"])==(w.end_i,w.end_j) and (it["start_i"],it["start_j"])==(w.start_i,w.start_j)
        if not ok:
            bad+=1
            if bad<=4: print("  BAD k=%d m=%d n=%d got=%s want=(%d,%d,%d,%d) len %d/%d"%(k,len(a),len(b),tuple(int(it[f]) for f in ("t1","t2","t3","end_state","score","end_i","end_j","start_i","start_j")),w.t1,w.t2,w.t3,w.end_state,it["aln_len"],len(w.ops)))
    print((g,h,mode,alpha,maxm,maxn),"bad",bad,"of",len(pairs))
