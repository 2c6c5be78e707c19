import numpy as np, sys
rng=np.random.default_rng(1)
with open(sys.argv[1],'w') as f:
    for s in range(int(sys.argv[2])):
        L=int(rng.integers(60,100))
        f.write(f">seq{s}\n"+"".join("ACGU"[b] for b in rng.integers(0,4,size=L))+"\n")
