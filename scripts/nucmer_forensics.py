#!/usr/bin/env python
"""What nucmer's own alignments say about its rules (analysis tool behind DESIGN.md section 4, "How the specification was pinned").

Works on the reference's checked-in nucmer 3.1 output (copied to oracle/_ref/fixtures by tests/golden/make_golden.py):
  scoring   every alignment is walked through its indel list and scored under candidate (gap open, gap extend) pairs; the
            right pair makes nucmer's path score equal the optimum of an unbanded global alignment between its end points
  dips      depth of every dip of the running score inside an alignment (high score -> lowest point before the next high
            score) and the length of the drought in anti-diagonals: what an extension survives
  gaps      distance and box score between consecutive colinear alignments that nucmer did NOT join
Usage: python scripts/nucmer_forensics.py [scoring|dips|gaps] [ref genome] [query genome]   (Caulobacter fixture names)
"""
import ctypes
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle, numpy as np
FIX = ROOT / 'oracle' / '_ref' / 'fixtures'
L=oracle.lib()
L.orc_global_score.restype=ctypes.c_int64
L.orc_global_score.argtypes=[ctypes.c_void_p,ctypes.c_void_p,ctypes.c_int64,ctypes.c_int64]+[ctypes.c_int]*4
for f in ('orc_genome_fw','orc_genome_rc'): getattr(L,f).restype=ctypes.c_void_p; getattr(L,f).argtypes=[ctypes.c_void_p]
L.orc_genome_npos.restype=ctypes.c_int64; L.orc_genome_npos.argtypes=[ctypes.c_void_p]
L.orc_genome_goff.restype=ctypes.c_uint64; L.orc_genome_goff.argtypes=[ctypes.c_void_p,ctypes.c_uint32]
def parse_delta_full(path):
    """-> list of (rid,qid,rs,re,qs,qe,errs,[deltas])"""
    out=[]; lines=open(path).read().split('\n'); i=2; ids=None
    while i<len(lines):
        l=lines[i]
        if l.startswith('>'): ids=l[1:].split()[:2]; i+=1; continue
        f=l.split()
        if len(f)==7:
            ds=[]; i+=1
            while lines[i]!='0': ds.append(int(lines[i])); i+=1
            out.append((ids[0],ids[1],int(f[0]),int(f[1]),int(f[2]),int(f[3]),int(f[4]),ds))
        i+=1
    return out
def path_stats(R,Q,ds):
    """R,Q: numpy code arrays of the aligned substrings (Q already oriented). returns matches, mismatches, list of gap run lengths"""
    i=j=0; m=x=0; runs=[]; cur=None
    for d in ds:
        adv=abs(d)-1
        a=R[i:i+adv]; b=Q[j:j+adv]
        eq=(a==b)&(a<4); m+=int(eq.sum()); x+=adv-int(eq.sum())
        i+=adv; j+=adv
        t=d>0
        if cur is not None and cur[0]==t and adv==0: cur[1]+=1
        else:
            cur=[t,1]; runs.append(cur)
        if d>0: i+=1
        else: j+=1
    a=R[i:]; b=Q[j:]
    assert len(a)==len(b),(len(a),len(b))
    eq=(a==b)&(a<4); m+=int(eq.sum()); x+=len(a)-int(eq.sum())
    return m,x,[r[1] for r in runs]

def codes(g, strand):
    n=L.orc_genome_npos(g._h)
    p=L.orc_genome_rc(g._h) if strand else L.orc_genome_fw(g._h)
    return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(n,))
def run(sub, ddir, r, q, fn=None, maxlen=3000, limit=400):
    g={}
    ids={}
    for n in (r,q):
        a,off,idl=oracle.read_fasta(FIX/sub/f'{n}.fna'); g[n]=oracle.OGenome(a,off); ids[n]=(idl,off)
    alns=parse_delta_full(FIX/ddir/(fn or f'{r}_vs_{q}.delta'))
    Rf=codes(g[r],0); Qf=codes(g[q],0); Qr=codes(g[q],1); nq=len(Qf)
    out=[]
    for (rid,qid,rs,re,qs,qe,errs,ds) in alns:
        if re-rs+1>maxlen or re-rs+1<100: continue
        rc=ids[r][0].index(rid); qc=ids[q][0].index(qid)
        ro=int(ids[r][1][rc])+rc; qo=int(ids[q][1][qc])+qc
        R=Rf[ro+rs-1:ro+re]
        if qs<=qe: Qv=Qf[qo+qs-1:qo+qe]
        else:
            # reverse strand: forward interval [qe,qs] -> rc view
            a=qo+qe-1; b=qo+qs  # forward [a,b)
            Qv=Qr[nq-b:nq-a]
        m,x,runs=path_stats(R,Qv,ds)
        assert x+sum(runs)==errs,(x,sum(runs),errs)
        out.append((R.copy(),Qv.copy(),m,x,runs,ds,(rs,re,qs,qe,errs)))
        if len(out)>=limit: break
    return out


def path_scores(R,Q,ds):
    """per-column (di,dj,score) along the nucmer path"""
    cols=[]
    i=j=0; prev=None
    def addmatch(n):
        nonlocal i,j,prev
        a=R[i:i+n]; b=Q[j:j+n]
        eq=(a==b)&(a<4)
        for e in eq: cols.append((1,1,3 if e else -7))
        i+=n; j+=n
        if n: prev=None
    for d in ds:
        addmatch(abs(d)-1)
        t=d>0
        cols.append((1,0,-(7 if prev==t else 10)) if t else (0,1,-(7 if prev==t else 10)))
        prev=t
        if t: i+=1
        else: j+=1
    addmatch(len(R)-i)
    return cols

def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "scoring"
    r = sys.argv[2] if len(sys.argv) > 2 else "NC_002696"
    q = sys.argv[3] if len(sys.argv) > 3 else "NC_010338"
    if what == "scoring":
        S = run("sequences", "anim_deltadir", r, q, limit=300)
        print(len(S), "alignments of", r, "vs", q)
        for go, ge in [(10, 7), (10, 4), (7, 7), (7, 4), (10, 3), (11, 4), (10, 5), (8, 4), (10, 2), (6, 3), (7, 3), (10, 1), (4, 4), (7, 1)]:
            worse = eq = 0
            for R, Qv, m, x, runs, _ds, _h in S:
                ps = 3 * m - 7 * x - sum(go + (k - 1) * ge for k in runs)
                opt = L.orc_global_score(R.ctypes.data, Qv.ctypes.data, len(R), len(Qv), 3, 7, go, ge)
                eq += ps == opt
                worse += ps < opt
            print((go, ge), "nucmer's path is optimal on", eq, "sub-optimal on", worse)
    elif what == "dips":
        S = run("sequences", "anim_deltadir", r, q, limit=100000, maxlen=1000000)
        depth, drought = [], []
        for R, Qv, m, x, runs, ds, h in S:
            s = hi = mn = ad = last = 0
            for di, dj, sc in path_scores(R, Qv, ds):
                s += sc; ad += di + dj
                if s >= hi:
                    if hi - mn > 0:
                        depth.append(hi - mn); drought.append(ad - last)
                    hi = mn = s; last = ad
                elif s < mn:
                    mn = s
        depth, drought = np.array(depth), np.array(drought)
        print(len(depth), "dips; depth max", depth.max(), "histogram", np.histogram(depth, bins=[0, 100, 200, 300, 400, 500, 550, 600, 650, 1000])[0].tolist())
        print("droughts (anti-diagonals) max", drought.max(), "over 200:", int((drought > 200).sum()))
    elif what == "gaps":
        g = {}
        for n in (r, q):
            a, off, _ = oracle.read_fasta(FIX / "sequences" / f"{n}.fna"); g[n] = oracle.OGenome(a, off)
        Rf, Qf, Qr = codes(g[r], 0), codes(g[q], 0), codes(g[q], 1)
        nq = len(Qf)
        al = []
        for rc, qc, rs, re, qs, qe, errs in oracle.parse_delta_text(FIX / "anim_deltadir" / f"{r}_vs_{q}.delta").tolist():
            al.append((0, rs - 1, re, qs - 1, qe) if qs <= qe else (1, rs - 1, re, nq - 1 - (qs - 1), nq - 1 - (qe - 1) + 1))
        al.sort()
        out = []
        for i, (st, rs, re, qs, qe) in enumerate(al):
            for st2, rs2, re2, qs2, qe2 in al[i + 1:i + 6]:
                if st2 != st:
                    break
                gr, gq = rs2 - re, qs2 - qe
                if gr < 0 or gq < 0 or gr > 400 or gq > 400:
                    continue
                Qs = Qr if st else Qf
                R, Q = Rf[re:rs2].copy(), Qs[qe:qs2].copy()
                sc = L.orc_global_score(R.ctypes.data, Q.ctypes.data, len(R), len(Q), 3, 7, 10, 7) if len(R) and len(Q) else -(10 + 7 * (max(len(R), len(Q)) - 1))
                out.append((gr, gq, sc))
                break
        out = np.array(out)
        print(len(out), "close colinear neighbours nucmer did not join; smallest max(gap_r, gap_q):", np.sort(np.maximum(out[:, 0], out[:, 1]))[:10].tolist())
        small = out[(out[:, 0] + out[:, 1]) <= 200]
        print("with gap_r + gap_q <= 200:", small.tolist())


if __name__ == "__main__":
    main()
