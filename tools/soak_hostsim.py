import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import random, numpy as np, sys, time
import helpers as H, oracle
from orcdemux import synth
import test_hostsim as T
seed=int(sys.argv[1]); trials=int(sys.argv[2])
iupac = len(sys.argv) > 3 and sys.argv[3] == "iupac"     # every adapter gets IUPAC wildcards (next row N4)
rnd = random.Random(seed)
t0=time.time(); total=0
for trial in range(trials):
    nf, nb = rnd.randint(1, 32), rnd.randint(1, 32)
    long_only = rnd.random() < 0.45       # long adapters at low error rates: stage 1 goes through the seed table
    lens = [40, 48, 57, 59, 64, 64] if long_only else [3, 5, 8, 12, 17, 20, 33, 40, 57, 64]
    mk = lambda: "".join(rnd.choice("ACGT") for _ in range(rnd.choice(lens)))
    shared = mk()[:rnd.choice([4, 10, 17, 25, 32, 40])]
    pshare = rnd.choice([0.5, 1.0, 1.0])
    sfx = mk()[:rnd.choice([0, 3, 8, 17, 30])]
    f = [(((shared if rnd.random() < pshare else "") + mk())[:64 - len(sfx)] + (sfx if rnd.random() < pshare else "")) for _ in range(nf)]
    sfx_b = mk()[:rnd.choice([0, 6, 13, 23, 30])]
    psfx = rnd.choice([0.5, 1.0, 1.0])
    b = [(((shared if rnd.random() < pshare else "") + mk())[:64 - len(sfx_b)] + (sfx_b if rnd.random() < psfx else "")) for _ in range(nb)]
    if rnd.random()<0.3:  # low complexity adapters
        f=[ (x[:4]*16)[:len(x)] for x in f]; b=[(x[:3]*22)[:len(x)] for x in b]
    e = rnd.choice([0.0, 0.05, 0.1, 0.1, 0.2, 0.3, 0.4, 0.6, 0.9, 2])
    if long_only: e = rnd.choice([0.0, 0.03, 0.05, 0.08, 0.1, 0.1, 0.12, 3])
    if e >= 1 and min(len(x) for x in f + b) <= e: e = 0.25
    ov = rnd.choice([1, 2, 3, 3, 5, 8, 20]); rc = rnd.choice([0, 1, 1])
    rounds = [(f, oracle.FRONT, e, ov, rc), (b, oracle.BACK, e, ov, rc)]
    if rnd.random()<0.25: rounds=[rounds[1], rounds[0]]   # back first then front
    if rnd.random()<0.15: rounds=rounds[:1]
    fi, bi = f, b
    if iupac:
        def wild(x):
            x = list(x)
            for _ in range(rnd.randint(1, max(1, len(x) // 6))):
                x[rnd.randrange(len(x))] = rnd.choice("RYSWKMBDHVNNI")
            if rnd.random() < 0.25 and len(x) >= 33:
                a0 = rnd.randrange(len(x) - 17); x[a0:a0 + 17] = "N" * 17
            return "".join(x)
        f, b = [wild(x) for x in f], [wild(x) for x in b]
        rounds = [(f if r[1] == oracle.FRONT else b,) + r[1:] for r in rounds]
        fi, bi = T._instances(rnd, f), T._instances(rnd, b)
    rs = T._adversarial_reads(rnd, fi, bi, 300)
    # add very short reads
    recs=[rs.read(i) for i in range(rs.n_reads)]
    for i in range(100):
        L=rnd.randint(0,90); a=rnd.choice(fi+bi)
        s="".join(rnd.choice("ACGT") for _ in range(L))
        if rnd.random()<0.6 and len(a)>2:
            x=rnd.randint(0,len(a)-1); y=rnd.randint(x+1,len(a)); p=rnd.randint(0,max(0,L))
            s=s[:p]+a[x:y]+s[p:]
        recs.append(("s%d"%i, s, "I"*len(s)))
    rs=synth.from_records(recs)
    rec0, rec1, oseq, oqual, olen = None, None, None, None, None
    fmode = rnd.choice([0, 1, 2, 2, 2]) | rnd.choice([0, 0, 0, 4])      # bit 2: keep the flank scan
    indels = rnd.choice([1, 1, 1, 0])
    try:
        m0, m1, lo, ln, rcv, nt = H.run_hostsim(rounds, rs, fmode, indels=indels)
    except RuntimeError as ex:          # e.g. an absolute error count no smaller than an adapter's non-N length
        if "unsupported" not in str(ex): raise
        continue
    rec0, rec1, oseq, oqual, olen = H.run_oracle(rounds, rs, n_threads=8, indels=bool(indels))
    total+=rs.n_reads
    for name, a, bb in (("r1", rec0, m0), ("r2", rec1, m1)):
        if a is None: continue
        idx, nbad = H.diff_matches(a, bb)
        if nbad:
            i=int(idx[0]); print("MISMATCH seed",seed,"trial", trial, name, "fmode", fmode, "indels", indels, "e", e, "ov", ov, "rc", rc, "read", i, "nbad", nbad)
            print(" oracle", a[i]); print(" hostsim", bb[i]); print(" seq", rs.read(i)[1]); print(rounds)
            sys.exit(1)
    assert np.array_equal(olen, ln)
import ctypes
sd = (ctypes.c_uint64 * 2)()
H.hostsim().hostsim_seeded(sd)
rb = (ctypes.c_uint64 * 2)()
H.hostsim().hostsim_resolved(rb)
print("ok seed",seed,"reads",total,"seeded (read, round) passes",sd[0],sd[1],"resolver tasks band/wide",rb[0],rb[1],"time",time.time()-t0)
