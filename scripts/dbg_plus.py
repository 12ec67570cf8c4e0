import sys, numpy as np, torch, pathlib, tempfile
sys.path.insert(0,'/root/repo')
from tests import _golden as G
from tests.test_gpu_plus import make_kg, make_model, ref_loss
DEV='cuda:0'
for name in ('syn','umls'):
    fx=G.load(name); kg=make_kg(fx)
    for tag in G.plus_tags(fx):
        cfg=G.plus_cfg(fx,tag)
        if cfg['entity_feature']=='RotatE': continue
        m,cfg=make_model(fx,kg,tag,pathlib.Path(tempfile.mkdtemp()))
        for j in range(1):
            tri,target,etr=G.train_batch_inputs(fx,j)
            t=torch.from_numpy(tri).to(DEV)
            m.zero_grad()
            score,mask=m(t[:,0],t[:,1],etr.to(DEV))
            loss=ref_loss(score,mask,target.to(DEV),t[:,2]); loss.backward()
            print(name,tag,cfg['type'],cfg['aggregator'],cfg['entity_feature'],'loss',loss.item(),fx['%s_tb%d_loss'%(tag,j)])
            for pn,par in m.named_parameters():
                key='%s_tb%d_g_%s'%(tag,j,pn)
                if key in fx and par.grad is not None:
                    w=fx[key]; g=par.grad.cpu().numpy()
                    print('   %-45s maxabs %.3e  err %.3e  rel %.3e'%(pn,np.abs(w).max(),np.abs(g-w).max(),np.abs(g-w).max()/max(1e-12,np.abs(w).max())))
