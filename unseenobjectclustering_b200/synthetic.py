"""Synthetic inputs of the BASELINE.json configs (no datasets or checkpoints exist offline)."""
import torch
import torch.nn.functional as F


def rgbd_frame(H=480, W=640, seed=0, batch=1):
    """image ~ U(-0.5, 0.6) (BGR/255 - PIXEL_MEANS/255 range), XYZ from Z ~ U(0.3, 1.5) m with the demo
    intrinsics of data/demo/camera_params.json scaled to HxW.  Returns ([B,3,H,W], [B,3,H,W]) float32 CPU."""
    g = torch.Generator().manual_seed(int(seed))
    img = torch.rand(batch, 3, H, W, generator=g) * 1.1 - 0.5
    z = torch.rand(batch, 1, H, W, generator=g) * 1.2 + 0.3
    sx, sy = W / 640.0, H / 480.0
    fx, fy, cx, cy = 612.937 * sx, 613.173 * sy, 322.549 * sx, 248.158 * sy
    u = torch.arange(W, dtype=torch.float32).view(1, 1, 1, W)
    v = torch.arange(H, dtype=torch.float32).view(1, 1, H, 1)
    xyz = torch.cat([(u - cx) / fx * z, (v - cy) / fy * z, z], dim=1)
    return img, xyz


def clustered_features(H, W, d=64, num_objects=6, noise=0.05, seed=0):
    """Unit-norm embedding field [1,d,H,W]: num_objects rectangles on a background,
    x = normalize(centre[label] + noise * randn).  Returns (features, gt [H,W])."""
    g = torch.Generator().manual_seed(int(seed))
    centres = F.normalize(torch.randn(num_objects + 1, d, generator=g), dim=1)
    gt = torch.zeros(H, W, dtype=torch.long)
    for k in range(1, num_objects + 1):
        h = int(torch.randint(H // 8, H // 3, (1,), generator=g))
        w = int(torch.randint(W // 8, W // 3, (1,), generator=g))
        y = int(torch.randint(0, H - h, (1,), generator=g))
        x = int(torch.randint(0, W - w, (1,), generator=g))
        gt[y:y + h, x:x + w] = k
    X = F.normalize(centres[gt.view(-1)] + noise * torch.randn(H * W, d, generator=g), dim=1)
    return X.t().contiguous().view(1, d, H, W), gt
