"""Seeded synthetic inputs of SURVEY.md section 8(d) (CPU generator => identical on
every machine with the same torch): plain torch, no oracle or product code.  Shared by
oracle/validate_against_reference.py, tests/, tools/ and bench.py."""
import torch


def synth_rpn(B, A, H, W, seed):
    """SURVEY.md 8(d) synthetic RPN outputs (CPU generator, deterministic)."""
    g = torch.Generator().manual_seed(seed)
    logits = 2.0 * torch.randn(B, 2, A * H, W, generator=g)
    prob = torch.softmax(logits, 1).view(B, 2 * A, H, W).contiguous()
    deltas = 0.2 * torch.randn(B, 4 * A, H, W, generator=g)
    return prob, deltas


def synth_gt(B, n_gt, K, seed, im_h=600, im_w=1200):
    g = torch.Generator().manual_seed(seed)
    gt = torch.zeros(B, K, 5)
    for b in range(B):
        x1 = torch.rand(n_gt, generator=g) * 1000
        y1 = torch.rand(n_gt, generator=g) * 450
        w = 20 + torch.rand(n_gt, generator=g) * 280
        h = 20 + torch.rand(n_gt, generator=g) * 180
        gt[b, :n_gt, 0] = x1
        gt[b, :n_gt, 1] = y1
        gt[b, :n_gt, 2] = torch.clamp(x1 + w, max=im_w - 1)
        gt[b, :n_gt, 3] = torch.clamp(y1 + h, max=im_h - 1)
        gt[b, :n_gt, 4] = torch.randint(1, 9, (n_gt,), generator=g).float()
    return gt


def synth_rois(R, B, seed, im_h=600, im_w=1200):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.rand(R, generator=g) * 1100
    y1 = torch.rand(R, generator=g) * 500
    w = 16 + torch.rand(R, generator=g) * 384
    h = 16 + torch.rand(R, generator=g) * 284
    b = torch.randint(0, B, (R,), generator=g).float()
    rois = torch.stack([b, x1, y1, torch.clamp(x1 + w, max=im_w - 1), torch.clamp(y1 + h, max=im_h - 1)], 1)
    return rois
