import sys, torch, json
sys.path.insert(0, "/root/repo")
from fmdm_b200 import ops
from fmdm_b200.models.vae import AutoencoderKL
from fmdm_b200.models.generators import DiffusionUNetFactory
from bench import LDCT_UNET
DEV = torch.device("cuda")
torch.manual_seed(0)
vae = AutoencoderKL(in_channels=1, out_channels=1, resolution=256, down_channels=(128, 256, 512, 512), num_res_blocks=2, z_channels=4, embed_dim=4, attn_heads=4, attn_dim_head=64)
for p in vae.parameters():
    if float(p.detach().abs().sum()) == 0: torch.nn.init.normal_(p, 0, 0.02)
vae = vae.to(DEV).eval()
B = 128
z = torch.randn(B, 4, 64, 64, device=DEV)
def prof(fn):
    with torch.no_grad():
        fn(); fn()
        with ops.profile() as rec:
            fn()
    by = {}
    for tag, work, ms in rec.rows:
        a = by.setdefault(tag, [0.0, 0.0, 0]); a[0] += work; a[1] += ms; a[2] += 1
    return {k: (round(v[1], 3), v[2], round(v[0] / v[1] / 1e9, 1) if v[1] else 0) for k, v in by.items()}
print("decode", json.dumps(prof(lambda: vae.decode(z, denorm=True))))
unet = DiffusionUNetFactory().build(dict(LDCT_UNET, in_channels=4, out_channels=4), "concatenate", 4).to(DEV).eval()
x = torch.randn(B, 4, 64, 64, device=DEV); c = torch.randn(B, 4, 64, 64, device=DEV); t = torch.full((B,), 500.0, device=DEV)
print("latent unet fwd", json.dumps(prof(lambda: unet(x, t, context=c))))
