import torch
M=N=K=8192
A=torch.randn(M,K,device='cuda',dtype=torch.bfloat16); B=torch.randn(N,K,device='cuda',dtype=torch.bfloat16); C=torch.empty(M,N,device='cuda',dtype=torch.bfloat16)
for _ in range(4): torch.matmul(A,B.t(),out=C)
A2=torch.randn(30720,4096,device='cuda',dtype=torch.bfloat16); B2=torch.randn(11008,4096,device='cuda',dtype=torch.bfloat16); C2=torch.empty(30720,11008,device='cuda',dtype=torch.bfloat16)
for _ in range(4): torch.matmul(A2,B2.t(),out=C2)
torch.cuda.synchronize(); print('done')
