import sys, types, numpy as np, torch, time
sys.path.insert(0,'/root/repo')
import bench
from volumetricrestirrelease_b200 import VolumetricReSTIR, VolumetricReSTIRParams
args=types.SimpleNamespace(kind='bunny',mips=4,dim=[577,572,438])
sc=bench.build_scene(args)
W,H=1920,1080
gp=VolumetricReSTIR.create({"mParams":VolumetricReSTIRParams()}); gp.setScene(sc,W,H)
color=torch.zeros((H,W,4),dtype=torch.float32,device='cuda')
for st in range(7):
    if st==3: gp.execute_stage(3,0,color.data_ptr())
    else: gp.execute_stage(st,0,color.data_ptr())
    torch.cuda.synchronize()
    rays,n=gp.debug_long_rays()
    print("stage",st,"long rays",n)
    for r in rays[:6]: print("   ", [float(x) for x in r])
np.save('gpurun_out/longrays.npy', rays)
print(gp.timings())
