import os, sys
sys.path.insert(0, '/root/repo')
exec(open('/root/repo/scripts/dev_stages.py').read().split("prof = eng.qp_profile(n)")[0])
prof = eng.qp_profile(n)
order = np.argsort(iters)
for a in order[-6:]:
    nf = max(1, prof[a,3])
    print(int(iters[a]), "nfact", int(prof[a,3]), "fact total", int(prof[a,1]), "per fact: assembly", int(prof[a,4]/nf), "cholesky", int(prof[a,5]/nf), "subst+tiles", int(prof[a,6]/nf), "setup(incl 1st fact)", int(prof[a,0]))
