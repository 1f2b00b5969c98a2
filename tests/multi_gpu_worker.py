"""Worker of tests/test_gpu_multi.py: launched by torchrun with one process per GPU (NCCL).  Every rank loads only the
raster rows of its own band (setOwnedRows -> halo rows by NCCL send / recv), runs its share of the path, the disjoint
output bands are gathered on rank 0 and compared bit for bit with a single-process run of the same engine -- in the
tile-by-tile mode and in dedup mode (accumulator seam strips by NCCL send / recv)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import golden_inputs
    import toy_models
    import moonsuperresolution_b200 as msr
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    dist.init_process_group("nccl", device_id=dev)
    case = dict(H=2600, W=1100, I=64, S=16, B=16, T=512, NV=-32768.0, seed=21, holes=True)
    dem, img = golden_inputs.make_rasters(case)
    ok = True
    for mode, model in (("faithful", toy_models.ripple), ("dedup", toy_models.ripple),
                        ("dedup", msr.IdentityModel(case["I"], case["B"]))):
        cfg = msr.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                            no_value=case["NV"], mode=mode)
        eng = msr.DEMSuperResolution(cfg, model=model, rank=rank, world_size=world, device=dev)
        o0, o1 = eng.ownedRows(case["H"], case["W"])
        eng.setOwnedRows(torch.from_numpy(dem[o0:o1]).to(dev), torch.from_numpy(img[o0:o1]).to(dev), case["H"])
        eng.padInputs()
        eng.processTiles()
        res = eng.gatherResults()
        if rank == 0:
            single = msr.DEMSuperResolution(cfg, model=model, device=dev).run(dem, img)
            same = all(np.array_equal(a, b) for a, b in zip(res, single))
            print(f"mode={mode} model={getattr(model, '__name__', type(model).__name__)} world={world}: "
                  f"{'bit-identical to the single-process run' if same else 'MISMATCH'}; good pixels {int(res[2].sum())}",
                  flush=True)
            ok = ok and same and int(res[2].sum()) > 0
        dist.barrier()
    # from files: every rank decodes only the strips of the rows its band reads (geotiff.read(rows=...))
    import tempfile
    from moonsuperresolution_b200 import geotiff
    box = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0)
    folder = box[0]
    if rank == 0:
        geotiff.write(os.path.join(folder, "run-DEM.tif"), dem, rows_per_strip=64)
        geotiff.write(os.path.join(folder, "run-DRG.tif"), img, rows_per_strip=64)
    dist.barrier()
    for mode in ("faithful", "dedup"):
        cfg = msr.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                            no_value=case["NV"], mode=mode, preprocess=False, source_folder_path=folder, map_name="m",
                            save_path=folder)
        eng = msr.DEMSuperResolution(cfg, model=toy_models.ripple, rank=rank, world_size=world, device=dev)
        eng.loadImages()
        held = eng.dem.shape[0]
        eng.padInputs()
        eng.processTiles()
        res = eng.gatherResults()
        if rank == 0:
            cfg1 = msr.DSRConfig(image_size=case["I"], stride=case["S"], batch_size=case["B"], tile_size=case["T"],
                                 no_value=case["NV"], mode=mode)
            single = msr.DEMSuperResolution(cfg1, model=toy_models.ripple, device=dev).run(dem, img)
            same = all(np.array_equal(a, b) for a, b in zip(res, single)) and held < case["H"]
            print(f"mode={mode} from files, {held} of {case['H']} rows decoded on rank 0, world={world}: "
                  f"{'bit-identical to the single-process run' if same else 'MISMATCH'}", flush=True)
            ok = ok and same
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
