import sys, numpy as np
sys.path.insert(0,'tests'); sys.path.insert(0,'orb-slam2-optimized_b200')
import oracle_api as O
from ransac_b200 import synth, capi
eng=capi.Engine(0)
C,n=8,1000
b=synth.pnp_batch(2,C,n,0.5)
cov=np.stack([synth.bearing_covariances(dict(K=b['K'],sigma2=b['sigma2'][c])) for c in range(C)])
Kf=np.array([b['K']],np.float32)
prm=dict(prob=0.99,min_inliers=10,max_its=300,min_set=6,eps=0.2,th2=5.991)
offsets=(np.arange(C+1)*n).astype(np.int32)
for use_cov in (True,False):
    res,masks=eng.mlpnp_solve(offsets,b['p3d'],b['p2d'],b['sigma2'],Kf,capi.ransac_params(**prm),cov=cov if use_cov else None,seeds=b['seeds'])
    poses,counts=eng.mlpnp_hypotheses()
    ml=eng.split_masks(masks,offsets)
    tot_bad=0
    for c in range(C):
        pb=O.mlpnp_problem(b['p3d'][c],b['p2d'][c],b['sigma2'][c],tuple(Kf[0]),cov[c] if use_cov else None)
        tab=O.index_table(int(b['seeds'][c]),n,6,300)
        o=O.mlpnp_ransac(pb,O.params(**prm),tab,O.FLAG_EXHAUSTIVE,per_hyp=True)
        gp=poses[c*300:(c+1)*300]; op=o['hyp_pose']
        d=np.abs(gp-op)/np.maximum(1,np.abs(op))
        dmax=np.nanmax(d,axis=1)
        bits=(gp.view(np.uint64)==op.view(np.uint64)).all(axis=1).sum()
        cd=(counts[c*300:(c+1)*300]!=o['hyp_counts'])
        print(c,'bit-identical hyps',bits,'max rel diff',np.nanmax(dmax),'>1e-9:',(dmax>1e-9).sum(),'count diffs',cd.sum(), 'res ok',res[c]['ok'],o['ok'],'ninl',res[c]['n_inliers'],o['n_inliers'],'besth',res[c]['best_hyp'],o['best_hyp'],'maskdiff',(ml[c]!=o['mask']).sum(),'dT',np.abs(res[c]['R'].reshape(3,3)-o['T'][:3,:3]).max(),np.abs(res[c]['t']-o['T'][:3,3]).max())
