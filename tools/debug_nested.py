import sys; sys.path.insert(0,'.')
import numpy as np
from newmsm_b200 import synth, resampler as R, build
from oracle import bindings as O
build.build_library()
for lo,hi in ((4,5),(3,5),(5,5),(4,6)):
    xyz,tri=synth.icosphere(lo); q,_=synth.icosphere(hi)
    rt,rv,rs=O.RefOctree(O.RefMesh(xyz,tri)).query(q,nthreads=8)
    ot,ov,os_,path=O.OracleOctree(xyz,tri).query(q)
    m=R.Mesh(xyz,tri); t=R.Octree(m)
    try:
        gt,gv,gs=t.query(q)
    except Exception as e:
        print('gpu exc',e); continue
    print(lo,hi,'ref vs oracle tri',(rt!=ot).sum(),'ref vs gpu tri',(rt!=gt).sum(),'status ref',np.bincount(rs),'gpu',np.bincount(gs), 'paths',np.bincount(path))
    bad=np.nonzero(rt!=gt)[0][:5]
    for b in bad: print('  q',b,q[b],'ref',rt[b],'gpu',gt[b],'path',path[b])
