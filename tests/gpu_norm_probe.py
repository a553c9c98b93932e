"""One-off probe (not a test): which float formula reproduces torch-CUDA's x.norm(dim=-1) on 3-vectors bit-for-bit?"""
import torch
torch.manual_seed(0)
x = torch.randn(1 << 20, 3, device='cuda')
ref = x.norm(dim=-1)
a, b, c = x[:, 0], x[:, 1], x[:, 2]
d = x.double()
A, B, C = d[:, 0], d[:, 1], d[:, 2]
f32 = lambda t: t.float().double()
fma = lambda p, q, r: f32(p * q + r)          # exact product in double, one rounding to double then float
cands = {
    'sqrt((aa+cc)+bb)': ((a * a + c * c) + b * b).sqrt(),
    'sqrt(aa+(bb+cc))': (a * a + (b * b + c * c)).sqrt(),
    'sqrt((bb+cc)+aa) ': ((b * b + c * c) + a * a).sqrt(),
    'sqrt(fma(b,b,fma(a,a,c*c)))': fma(B, B, fma(A, A, f32(C * C))).sqrt().float(),
    'sqrt(fma(a,a,c*c)+b*b)': f32(fma(A, A, f32(C * C)) + f32(B * B)).sqrt().float(),
    'sqrt(fma(c,c,a*a)+b*b)': f32(fma(C, C, f32(A * A)) + f32(B * B)).sqrt().float(),
    'norm via vector_norm f64 acc': torch.linalg.vector_norm(x, dim=-1, dtype=torch.float64).float(),
    'sqrt((aa+bb)+cc) no fma': ((a * a + b * b) + c * c).sqrt(),
    'sqrt(fma(c,c,fma(b,b,a*a)))': fma(C, C, fma(B, B, f32(A * A))).sqrt().float(),
    'sqrt(fma(a,a,fma(b,b,c*c)))': fma(A, A, fma(B, B, f32(C * C))).sqrt().float(),
    'float64 then round': (A * A + B * B + C * C).sqrt().float(),
    'sqrt(fma(c,c,fma(b,b,fma(a,a,0))))': fma(C, C, fma(B, B, fma(A, A, torch.zeros_like(A)))).sqrt().float(),
}
for k, v in cands.items():
    print(f'{k:45s} mismatches: {int((v != ref).sum())} / {ref.numel()}')
y = torch.randn(1 << 20, 3, device='cuda')
print('abs().amax exact:', bool((y.abs().amax(-1) == torch.maximum(torch.maximum(y[:, 0].abs(), y[:, 1].abs()), y[:, 2].abs())).all()))
