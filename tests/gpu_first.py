import sys, time, numpy as np
sys.path.insert(0,'.')
from smoothxg_b200 import engine, synth
from oracle.oracle import Oracle, make_params as op
from tests.helpers import view_to_dump, first_diff
ora=Oracle()
def run(name, batch, kw, **eng_kw):
    eng=engine.PoaEngine(device=0, emit_cigar=True, **eng_kw)
    t=time.time(); res=eng.run_batch(batch, engine.make_params(**kw), allow_block_errors=True); dt=time.time()-t
    bad=0
    for b in range(batch.n_blocks):
        v=res.block(b)
        if v.status!=0: print(name,b,"status",v.status); bad+=1; continue
        want=ora.poa_block(op(**kw),*batch.block(b)); got=view_to_dump(v)
        if not np.array_equal(want.compare_part(), got.compare_part()):
            bad+=1; print(name,b,first_diff(want,got))
    st=res.stats()
    print(name, "OK" if not bad else f"BAD {bad}/{batch.n_blocks}", f"{dt:.3f}s", {k:st[k] for k in ('kernel_ms','n_ctas','warps_per_block','kernel_launches','retried_blocks','inband_cells')}, st['phase_cycles'], flush=True)
    eng.close()
tiny=synth.PoaBatch.from_strings([["ACGTACGT","ACGTTCGT","ACGACGT"],["A"],["ACGT","ACGT"],["AC","","G"],[]])
for nw in (1,2,4,8):
    run(f"tiny nw{nw}", tiny, dict(), warps_per_block=nw)
    run(f"g8 nw{nw}", synth.make_batch(n_blocks=6,n_seqs=8,length=700,seed=3), dict(), warps_per_block=nw)
    run(f"local nw{nw}", synth.make_batch(n_blocks=4,n_seqs=6,length=500,seed=4), dict(local=True,out_msa=True), warps_per_block=nw)
    run(f"indel nw{nw}", synth.make_batch(n_blocks=4,n_seqs=8,length=1200,seed=5,indel_prob=0.5,n_frac=0.01,dup_weights=True), dict(out_msa=True), warps_per_block=nw)
    run(f"unb nw{nw}", synth.make_batch(n_blocks=4,n_seqs=6,length=400,seed=6,divergence=0.1), dict(banded=False), warps_per_block=nw)
run("retry", synth.make_batch(n_blocks=4,n_seqs=8,length=600,seed=9,divergence=0.2), dict(), slab_rows_factor=1.0)
