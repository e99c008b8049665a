import time, numpy as np, sys
sys.path.insert(0,'/root/repo')
from sawfish_b200 import synth
from sawfish_b200.batch import AlignContext, MultiGpuContext, CONSENSUS_PENALTIES
arena,pairs,_=synth.config2b_score_sv_flanks(200000)
ctx=AlignContext(0)
for _ in range(2): ctx.align_batch(CONSENSUS_PENALTIES, arena, pairs, False)
t=time.perf_counter()
for _ in range(3): s1=ctx.align_batch(CONSENSUS_PENALTIES, arena, pairs, False)[0]
dt=(time.perf_counter()-t)/3
print("single ctx: %.1f ms  %.2f M aln/s"%(dt*1e3, len(pairs)/dt/1e6), ctx.stats()["kernel_ms"])
for devs in ([0,0],[0,0,0]):
    m=MultiGpuContext(devs)
    for _ in range(2): m.align_batch(CONSENSUS_PENALTIES, arena, pairs, False)
    t=time.perf_counter()
    for _ in range(3): s2=m.align_batch(CONSENSUS_PENALTIES, arena, pairs, False)[0]
    dt=(time.perf_counter()-t)/3
    print(devs, "multi: %.1f ms  %.2f M aln/s"%(dt*1e3, len(pairs)/dt/1e6), m.last_kernel_ms(), (s1==s2).all())
    m.close()
