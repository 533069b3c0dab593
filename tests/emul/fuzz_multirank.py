"""tests/emul/fuzz_multirank.py — randomized differential test of block-decomposed solves: several ranks of the
host-emulated library (threads) against the SINGLE-BLOCK oracle, bit for bit.

    python tests/emul/fuzz_multirank.py loops|tma FIRST_SEED NUMBER_OF_CASES

Random block grids (2..8 ranks, all split patterns), ragged block sizes, walls, fixed / converged solves, NVLink
peer-memory or NCCL exchange, the three CG forms, the multigrid preconditioner (65 cases were clean when written).
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]
import build_emul
from cajitafluids_b200._capi import Library
from cajitafluids_b200 import config as K
from helpers import make_cfg, fields_of
from multirank import run_ranks, block_slices
from oracle_api import Oracle
which=sys.argv[1]; seed0=int(sys.argv[2]); ncase=int(sys.argv[3])
lib=Library(build_emul.build(tma=which=="tma"),"cfb_",mode=ctypes.RTLD_LOCAL)
grids=[(1,1,2),(1,2,1),(2,1,1),(1,2,2),(2,2,1),(2,1,2),(2,2,2),(3,1,1),(1,1,3),(1,3,2),(3,2,1),(4,1,1),(1,4,2)]
bad=0
for case in range(ncase):
    rng=np.random.default_rng(seed0+case)
    dim=int(rng.choice([2,3,3]))
    bl=grids[int(rng.integers(len(grids)))]
    if dim==2: bl=(bl[0],bl[1],1)
    world=bl[0]*bl[1]*bl[2]
    hi=40 if which=="loops" else 28
    cells=tuple(int(rng.integers(4*bl[d], max(4*bl[d]+1,hi))) for d in range(dim))
    mg = rng.random()<0.35
    if mg: cells=tuple(int(bl[d]*2*rng.integers(2,7)) for d in range(dim))
    h=2.0**-int(rng.integers(4,8)); box=tuple(c*h for c in cells)
    bt=[int(rng.choice([K.SOLID,K.SOLID,K.FREE])) for _ in range(2*dim)]
    fixed=int(rng.choice([0,9,17]))
    peer=bool(rng.integers(2)); variant=int(rng.choice([0,1,2]))
    rng2=np.random.default_rng(seed0+case+1000003)  # its own stream: earlier cases stay what they were
    flat=int(rng2.integers(2)); fusedx=int(rng2.integers(2))
    desc=f"case {seed0+case}: dim={dim} blocks={bl} cells={cells} bt={bt} fixed={fixed} peer={peer} variant={variant} mg={mg} flat_2d={flat} peer_overlap={fusedx}"
    try:
        cfg=make_cfg(dim,cells,box=box,boundary_type=bt,fixed_iters=fixed,max_iter=4000)
        ora=Oracle(cfg)
        if mg:
            n=[c//b for c,b in zip(cells,bl[:dim])]; lv=1
            while all(e%2==0 and e//2>=2 for e in n): n=[e//2 for e in n]; lv+=1
            ora.set_mg_max_levels(lv); ora.set_preconditioner("mg")
        vel={f:rng.uniform(-1,1,size=ora.shape(f)) for f in fields_of(dim)[1:]}
        for f,a in vel.items(): ora.set(f,a)
        ora.add_inputs(); ora.build_rhs()
        try: ro=ora.pcg_solve()
        except Exception as e: ro=("err",)
        po=ora.get(K.PRESSURE)
        def body(ctx,rank):
            ctx.set_tuning("cg_variant",variant)
            ctx.set_tuning("flat_2d",flat)
            ctx.set_tuning("peer_overlap",fusedx)
            if mg: ctx.set_tuning("mg_tma", 1); ctx.set_preconditioner("mg")
            for f,a in vel.items(): ctx.set(f,a[block_slices(ctx,f)])
            ctx.add_inputs(); ctx.build_rhs()
            try: rg=ctx.pcg_solve()
            except Exception as e: rg=("err",)
            return rg, bool(np.array_equal(ctx.get(K.PRESSURE),po[block_slices(ctx,K.PRESSURE)]))
        res=run_ranks(lib,cfg,world,body,bl,peer=peer,timeout=600)
        ok=all(r==(ro,True) or (ro==("err",) and r[0]==("err",)) for r in res)
        print(("ok  " if ok else "BAD ")+desc+f" -> {ro}",flush=True); bad+=(not ok)
    except Exception as e:
        print("EXC "+desc+" "+repr(e)[:200],flush=True); bad+=1
print("bad:", bad)
sys.exit(1 if bad else 0)
