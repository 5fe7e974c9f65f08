"""Builds oracle/_ref/ — the reference's OWN code, compiled / staged where it lies under /root/reference (build container
only; the outputs are git-ignored and travel to the GPU box with the snapshot).  TEST INFRASTRUCTURE, never the product.

* `libfps_ref.so`: the reference's furthest-point-sampling kernel, third_party/pointnet2/_ext_src/src/sampling_gpu.cu
  (:72-176 kernel, :178-215 host wrapper), compiled for sm_100a straight from the reference tree together with
  oracle/fps_ref_shim.cu (an extern "C" entry around the reference's own host wrapper).  It needs ATen only for
  `at::cuda::getCurrentCUDAStream()`, so it links against the torch libraries of this image.
* `libplanes_ref.so`: the reference's tri-plane projection functions, utils/cuda_utils/cuda_utils_kernel.cu (:371-600), compiled
  the same way with oracle/planes_ref_shim.cu (needs --expt-relaxed-constexpr: the file calls std::sqrt in a kernel; ~2 min).
* `libproject_ref.so`: the reference's 2D -> 3D feature projection, utils/cuda_utils/project_image_cuda_kernel.cu, likewise with
  oracle/project_ref_shim.cu.
* `felzenszwalb_ref/felzenszwalb_cpp*.so`: the reference's pybind module utils/cpp_utils/segmentator.cpp, compiled as it is.
* `reference/`: the UNMODIFIED reference Python files of the hot path (models/, pointnet2_utils.py, the pseudo-mask
  functions, the trainer and its entry point) staged so that the `-m gpu` tests can execute them on the CUDA shim on the GPU
  box, where /root/reference does not exist.  Nothing here is committed: the reference's sources stay out of the history.

    python oracle/build_ref.py [--verbose]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REFERENCE = "/root/reference"
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

STAGED = [
    "models", "third_party/pointnet2/pointnet2_utils.py", "third_party/pointnet2/pytorch_utils.py",
    "pseudo_masks/unscene3d_pseudo_main.py", "pseudo_masks/freemask_main.py", "utils/freemask_utils.py", "utils/pc_utils.py",
    "utils/utils.py", "utils/kfold.py", "utils/votenet_utils", "datasets/utils.py", "trainer/trainer.py", "trainer/__init__.py",
    "main_instance_segmentation.py", "conf", "benchmark", "models/metrics", "utils/point_cloud_utils.py", "datasets/scannet200",
    "datasets/__init__.py", "utils/__init__.py", "utils/cuda_utils/cuda_utils.py", "utils/cuda_utils/raycast_image.py",
    "models/noise_robust_loss.py",
]


def build_fps(verbose=False):
    import torch  # noqa: F401  (locates the headers / libraries)
    from torch.utils import cpp_extension as ce

    src = os.path.join(REFERENCE, "third_party/pointnet2/_ext_src/src/sampling_gpu.cu")
    inc = os.path.join(REFERENCE, "third_party/pointnet2/_ext_src/include")
    shim = os.path.join(HERE, "fps_ref_shim.cu")
    lib = os.path.join(OUT, "libfps_ref.so")
    if os.path.exists(lib) and os.path.getmtime(lib) > os.path.getmtime(shim):
        return lib
    tlib = ce.library_paths(device_type="cuda")[0]
    cmd = [NVCC, "-shared", "-Xcompiler", "-fPIC", "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-I", inc]
    for p in ce.include_paths(device_type="cuda"):
        cmd += ["-I", p]
    cmd += [src, shim, "-o", lib, "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-Xlinker", "-rpath," + tlib]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return lib


def build_planes(verbose=False, src_rel="utils/cuda_utils/cuda_utils_kernel.cu", shim_name="planes_ref_shim.cu", lib_name="libplanes_ref.so"):
    import sysconfig

    import torch  # noqa: F401
    from torch.utils import cpp_extension as ce

    src = os.path.join(REFERENCE, src_rel)
    shim = os.path.join(HERE, shim_name)
    lib = os.path.join(OUT, lib_name)
    if os.path.exists(lib) and os.path.getmtime(lib) > os.path.getmtime(shim):
        return lib
    tlib = ce.library_paths(device_type="cuda")[0]
    cmd = [NVCC, "-shared", "-Xcompiler", "-fPIC", "-O2", "-std=c++17", "--expt-relaxed-constexpr", "-w", "-gencode",
           "arch=compute_100a,code=sm_100a", "-I", os.path.join(REFERENCE, "utils/cuda_utils"), "-I", sysconfig.get_paths()["include"]]
    for p in ce.include_paths(device_type="cuda"):
        cmd += ["-I", p]
    cmd += [src, shim, "-o", lib, "-L", tlib, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-Xlinker", "-rpath," + tlib]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return lib


def build_felzenszwalb(verbose=False):
    """The reference's pybind module felzenszwalb_cpp, utils/cpp_utils/segmentator.cpp, compiled as it is (g++ -O2, the flags of
    a default setuptools build) into oracle/_ref/felzenszwalb_ref/felzenszwalb_cpp<ext suffix>: imported by the tests under that
    directory as the checker for us3d_felzenszwalb_segment_h."""
    import sysconfig

    import pybind11

    src = os.path.join(REFERENCE, "utils/cpp_utils/segmentator.cpp")
    out_dir = os.path.join(OUT, "felzenszwalb_ref")
    os.makedirs(out_dir, exist_ok=True)
    lib = os.path.join(out_dir, "felzenszwalb_cpp" + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(lib):
        return lib
    cmd = ["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-fvisibility=hidden", "-w", "-I", pybind11.get_include(), "-I",
           sysconfig.get_paths()["include"], "-I", os.path.join(REFERENCE, "utils/cpp_utils"), src, "-o", lib]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.run(cmd, check=True)
    return lib


def stage_reference():
    dst_root = os.path.join(OUT, "reference")
    for rel in STAGED:
        src = os.path.join(REFERENCE, rel)
        dst = os.path.join(dst_root, rel)
        if not os.path.exists(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.isdir(src):
            shutil.copytree(src, dst, dirs_exist_ok=True, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"), copy_function=shutil.copyfile)
        else:
            shutil.copyfile(src, dst)
    return dst_root


def build(verbose=False):
    """Returns None where /root/reference is absent — the GPU box uses what was built here."""
    if not os.path.isdir(REFERENCE):
        return None
    os.makedirs(OUT, exist_ok=True)
    # the two torch-extension libraries take ~5 minutes of nvcc each: all native builds run side by side; the GPU tests skip a
    # reference-kernel comparison whose library is absent
    from concurrent.futures import ThreadPoolExecutor

    jobs = {
        "fps": lambda: build_fps(verbose),
        "felzenszwalb": lambda: build_felzenszwalb(verbose),
        "planes": lambda: build_planes(verbose),
        "project": lambda: build_planes(verbose, "utils/cuda_utils/project_image_cuda_kernel.cu", "project_ref_shim.cu", "libproject_ref.so"),
    }
    out = [stage_reference()]
    with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        futures = {name: pool.submit(fn) for name, fn in jobs.items()}
        for name, fut in futures.items():
            try:
                out.append(fut.result())
            except Exception as e:  # noqa: BLE001
                print(f"oracle/_ref: {name} not built ({e})", file=sys.stderr)
    return tuple(out)


if __name__ == "__main__":
    print(build(verbose="--verbose" in sys.argv))
