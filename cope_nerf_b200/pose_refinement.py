"""Photometric relative-pose refinement between the two training stages (reference: utils_poses/pose_refinement.py:34-61
`compute_loss_and_warp_image`, :104-150 `perform_pose_refinement`) on libcope_b200.

The dense warp + masked L1 of one direction is ONE forward and ONE backward launch (cope_pose_refine_fwd / _bwd) instead of
the reference's ~40 ATen launches (batched 3x3 inverses and matmuls, grid_sample, masks); the gradient reaches the
`PoseRetriever` parameters (r, t) of every pair through cope_pose_bwd.  Dataset plumbing, logging and the pose metrics of the
reference's loop stay with the caller (out of scope: disk I/O and numpy metrics)."""
import torch

from . import _lib as L
from .losses import rigid_inverse

__all__ = ["compute_loss_and_warp_image", "make_uv", "refinement_losses", "perform_pose_refinement"]


def make_uv(resolution, device):
    """pose_refinement.py:88-96: the normalised pixel grid [3, H, W] = (2 col / (W-1) - 1, 2 row / (H-1) - 1, 1).
    The kernels rebuild it from (H, W); this is only for callers that keep the reference's argument list."""
    h, w = int(resolution[0]), int(resolution[1])
    rows, cols = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=device),
                                torch.arange(w, dtype=torch.float32, device=device), indexing="ij")
    return torch.stack([cols / ((w - 1) / 2) - 1, rows / ((h - 1) / 2) - 1, torch.ones_like(rows)])


class _RefineFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, images, next_images, depths, K, poses, want_warped):
        B, _, H, W = images.shape
        dev = images.device
        args = [t.contiguous().float() for t in (images, next_images, depths.reshape(B, H, W), K.reshape(B, 3, 3), poses)]
        warped = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev) if want_warped else None
        loss, ws = torch.empty(1, dtype=torch.float32, device=dev), torch.empty(4, dtype=torch.float32, device=dev)
        L.call("cope_pose_refine_fwd", *[L.ptr(a) for a in args], B, H, W, L.ptr(warped), L.ptr(loss), L.ptr(ws), L.stream())
        ctx.save_for_backward(*args, ws)
        ctx.dims = (B, H, W)
        ctx.set_materialize_grads(False)
        if warped is not None:
            ctx.mark_non_differentiable(warped)
        return loss.reshape(()), warped

    @staticmethod
    def backward(ctx, g, _unused=None):
        if g is None or not ctx.needs_input_grad[4]:
            return (None,) * 6
        *args, ws = ctx.saved_tensors
        B, H, W = ctx.dims
        d_poses = torch.zeros(B, 4, 4, dtype=torch.float32, device=ws.device)
        L.call("cope_pose_refine_bwd", *[L.ptr(a) for a in args], B, H, W, L.ptr(ws), L.ptr(g.reshape(1).float()), L.ptr(d_poses),
               L.stream())
        return None, None, None, None, d_poses, None


def compute_loss_and_warp_image(images, next_images, depths, K_batch, uv_batch=None, relative_poses=None, warp_pixel_fn=None,
                                return_warped=True):
    """pose_refinement.py:34-61, same argument order.  images / next_images [B,3,H,W], depths [B,1,H,W], K_batch [B,3,3],
    relative_poses [B,4,4].  `uv_batch` (the normalised pixel grid of :88-96) and `warp_pixel_fn` (train.py:235-244 with
    normalize_pix=False) are accepted for signature compatibility: the kernel builds the grid from (H, W) and samples with the
    same bilinear / border / align_corners=True rule.  Returns (loss, warped_images)."""
    return _RefineFn.apply(images, next_images, depths, K_batch, relative_poses, return_warped)


def refinement_losses(relative_pose_retriever, image_idx, images, next_images, depths, next_depths, K):
    """pose_refinement.py:117-126: the symmetric photometric loss of one batch of frame pairs — frame -> next frame with the
    pair's relative pose, next frame -> frame with its inverse, averaged.  Depths are [B,1,H,W] at the image resolution."""
    poses = torch.stack([relative_pose_retriever(int(i)) for i in image_idx])
    inv = torch.stack([rigid_inverse(p) for p in poses])        # torch.inverse of a rigid map (:124), without the LU
    l_pos, _ = compute_loss_and_warp_image(images, next_images, depths, K, None, poses, return_warped=False)
    l_neg, _ = compute_loss_and_warp_image(next_images, images, next_depths, K, None, inv, return_warped=False)
    return (l_pos + l_neg) / 2


def perform_pose_refinement(relative_pose_retriever, optimizer, batches, epochs, scheduler=None, resolution=None,
                            converge_std=1e-5, window=50):
    """The optimisation loop of pose_refinement.py:104-150 without its logging / pose-metric side: `batches` is any re-iterable
    of (image_idx, next_image_idx, images, next_images, depths, next_depths, K, next_K) tuples as PoseRefineDataset yields them
    (depths [B,h,w] are resized to `resolution` with nearest interpolation, :111-114).  Stops early when the running loss of
    the last `window` epochs has a standard deviation <= converge_std (:146-149).  Returns the list of per-epoch mean losses."""
    history = []
    for _ in range(epochs):
        running = torch.zeros((), dtype=torch.float32, device=relative_pose_retriever.r.device)
        count = 0
        for batch in batches:
            image_idx, _, images, next_images, depths, next_depths, K, _ = batch
            dev = relative_pose_retriever.r.device
            images, next_images, K = images.float().to(dev), next_images.float().to(dev), K.float().to(dev)
            depths, next_depths = depths.float().to(dev).unsqueeze(1), next_depths.float().to(dev).unsqueeze(1)
            if resolution is not None:
                size = (int(resolution[0]), int(resolution[1]))
                depths = torch.nn.functional.interpolate(depths, size)
                next_depths = torch.nn.functional.interpolate(next_depths, size)
            loss = refinement_losses(relative_pose_retriever, image_idx, images, next_images, depths, next_depths, K)
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            running += loss.detach() * len(images)      # no host sync inside the epoch (the reference calls .item() per batch)
            count += len(images)
        if scheduler is not None:
            scheduler.step()
        history.append(float(running) / max(count, 1))
        if len(history) >= window and torch.tensor(history[-window:]).std().item() <= converge_std:
            break
    return history
