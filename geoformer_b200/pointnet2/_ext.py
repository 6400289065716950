"""The nine pointnet2._ext operators (lib/pointnet2/_ext_src/src/bindings.cpp:9-22) on top of the
sm_100a C-ABI library.  Same names, argument order, dtypes, layouts, output allocation and error
behaviour as the reference's pybind module: float32 / int32 contiguous CUDA tensors in, freshly
allocated tensors out, RuntimeError on a violated precondition ("CPU not supported" included).
Kernels are queued on the current CUDA stream of the tensors' device and do not synchronise.
"""
import torch

from .. import _capi as C


import contextlib

_SAME_DEVICE = contextlib.nullcontext()


def _dev(t):
    # a device guard only when the tensor does not live on the current device (the guard costs more
    # than the launch of the small operators)
    return _SAME_DEVICE if t.device.index == torch.cuda.current_device() else torch.cuda.device(t.device)


def furthest_point_sampling(points, nsamples):
    """(B,N,3) f32 -> (B,nsamples) i32.  sampling.cpp:67-88."""
    C.check_cuda_f32(points, "points")
    C.require(points.dim() == 3 and points.size(2) == 3, "points must be (B, N, 3)")
    B, N, _ = points.shape
    nsamples = int(nsamples)
    out = torch.empty((B, nsamples), dtype=torch.int32, device=points.device)
    if B == 0 or nsamples <= 0:
        return out
    with _dev(points):
        L = C.lib()
        nbytes = L.gf_fps_workspace_bytes(B, N, nsamples)
        ws = C.workspace.get(points.device, "fps", nbytes) if nbytes else None
        C.check(L.gf_furthest_point_sampling(C.ptr(points), B, N, nsamples, C.ptr(out), C.ptr(ws), nbytes,
                                             C.stream_of(points.device)), "furthest_point_sampling")
    return out


def gather_points(points, idx):
    """(B,C,N) f32, (B,m) i32 -> (B,C,m).  sampling.cpp:17-40."""
    C.check_cuda_f32(points, "points")
    C.check_cuda_i32(idx, "idx")
    B, Cc, N = points.shape
    m = idx.size(1)
    out = torch.empty((B, Cc, m), dtype=torch.float32, device=points.device)
    with _dev(points):
        C.check(C.lib().gf_gather_points(C.ptr(points), C.ptr(idx), B, Cc, N, m, C.ptr(out),
                                         C.stream_of(points.device)), "gather_points")
    return out


def gather_points_grad(grad_out, idx, n):
    """(B,C,m) f32, (B,m) i32 -> (B,C,n).  sampling.cpp:42-65."""
    C.check_cuda_f32(grad_out, "grad_out")
    C.check_cuda_i32(idx, "idx")
    B, Cc, m = grad_out.shape
    out = torch.empty((B, Cc, int(n)), dtype=torch.float32, device=grad_out.device)
    with _dev(grad_out):
        C.check(C.lib().gf_gather_points_grad(C.ptr(grad_out), C.ptr(idx), B, Cc, int(n), m, C.ptr(out),
                                              C.stream_of(grad_out.device)), "gather_points_grad")
    return out


def ball_query(new_xyz, xyz, radius, nsample):
    """centres (B,m,3), points (B,N,3) -> (B,m,nsample) i32.  ball_query.cpp:11-35."""
    C.check_cuda_f32(new_xyz, "new_xyz")
    C.check_cuda_f32(xyz, "xyz")
    B, m, _ = new_xyz.shape
    N = xyz.size(1)
    nsample = int(nsample)
    out = torch.empty((B, m, nsample), dtype=torch.int32, device=new_xyz.device)
    with _dev(new_xyz):
        C.check(C.lib().gf_ball_query(C.ptr(new_xyz), C.ptr(xyz), B, N, m, float(radius), nsample, C.ptr(out),
                                      C.stream_of(new_xyz.device)), "ball_query")
    return out


def group_points(points, idx):
    """(B,C,N) f32, (B,np,ns) i32 -> (B,C,np,ns).  group_points.cpp:15-38."""
    C.check_cuda_f32(points, "points")
    C.check_cuda_i32(idx, "idx")
    B, Cc, N = points.shape
    _, np_, ns = idx.shape
    out = torch.empty((B, Cc, np_, ns), dtype=torch.float32, device=points.device)
    with _dev(points):
        C.check(C.lib().gf_group_points(C.ptr(points), C.ptr(idx), B, Cc, N, np_, ns, C.ptr(out),
                                        C.stream_of(points.device)), "group_points")
    return out


def group_points_grad(grad_out, idx, n):
    """(B,C,np,ns) f32, (B,np,ns) i32 -> (B,C,n).  group_points.cpp:40-62."""
    C.check_cuda_f32(grad_out, "grad_out")
    C.check_cuda_i32(idx, "idx")
    B, Cc, np_, ns = grad_out.shape
    out = torch.empty((B, Cc, int(n)), dtype=torch.float32, device=grad_out.device)
    with _dev(grad_out):
        C.check(C.lib().gf_group_points_grad(C.ptr(grad_out), C.ptr(idx), B, Cc, int(n), np_, ns, C.ptr(out),
                                             C.stream_of(grad_out.device)), "group_points_grad")
    return out


def three_nn(unknowns, knows):
    """(B,n,3), (B,m,3) -> [dist2 (B,n,3) f32, idx (B,n,3) i32].  interpolate.cpp:19-46."""
    C.check_cuda_f32(unknowns, "unknowns")
    C.check_cuda_f32(knows, "knows")
    B, n, _ = unknowns.shape
    m = knows.size(1)
    dist2 = torch.empty((B, n, 3), dtype=torch.float32, device=unknowns.device)
    idx = torch.empty((B, n, 3), dtype=torch.int32, device=unknowns.device)
    with _dev(unknowns):
        C.check(C.lib().gf_three_nn(C.ptr(unknowns), C.ptr(knows), B, n, m, C.ptr(dist2), C.ptr(idx),
                                    C.stream_of(unknowns.device)), "three_nn")
    return [dist2, idx]


def three_interpolate(points, idx, weight):
    """(B,c,m) f32, (B,n,3) i32, (B,n,3) f32 -> (B,c,n).  interpolate.cpp:48-79."""
    C.check_cuda_f32(points, "points")
    C.check_cuda_i32(idx, "idx")
    C.check_cuda_f32(weight, "weight")
    B, c, m = points.shape
    n = idx.size(1)
    out = torch.empty((B, c, n), dtype=torch.float32, device=points.device)
    with _dev(points):
        C.check(C.lib().gf_three_interpolate(C.ptr(points), C.ptr(idx), C.ptr(weight), B, c, m, n, C.ptr(out),
                                             C.stream_of(points.device)), "three_interpolate")
    return out


def three_interpolate_grad(grad_out, idx, weight, m):
    """(B,c,n) f32, (B,n,3) i32, (B,n,3) f32 -> (B,c,m).  interpolate.cpp:81-113."""
    C.check_cuda_f32(grad_out, "grad_out")
    C.check_cuda_i32(idx, "idx")
    C.check_cuda_f32(weight, "weight")
    B, c, n = grad_out.shape
    out = torch.empty((B, c, int(m)), dtype=torch.float32, device=grad_out.device)
    with _dev(grad_out):
        C.check(C.lib().gf_three_interpolate_grad(C.ptr(grad_out), C.ptr(idx), C.ptr(weight), B, c, n, int(m),
                                                  C.ptr(out), C.stream_of(grad_out.device)), "three_interpolate_grad")
    return out
