import sys
sys.path.insert(0,'/root/repo/vision-ft_b200'); sys.path.insert(0,'/root/repo')
import torch
from vft_b200 import _cabi
torch.zeros(1, device='cuda')
for (T,N,K) in [(528,3072,3072),(2,18432,3072),(154,640,2048),(256,3840,2304),(4096,3072,3072)]:
    print(T,N,K, 'fwd ws', _cabi.lib.vft_workspace_bytes(0,T,N,K,0), 'bwd ws', _cabi.lib.vft_workspace_bytes(1,T,N,K,0))
