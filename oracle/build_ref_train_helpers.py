"""TEST INFRASTRUCTURE.  Compiles the REFERENCE's own training-loop helpers to bytecode in oracle/_ref/train_helpers.bin.

The helpers are the functions of the three training scripts that call the hot-path modules:
    train.py:105-195        sample_predict_points, get_vp_meshes, compose_vp_meshes, calculate_cd_loss,
                            calculate_silhouette_loss, calculate_vp_div_loss, calculate_emd_loss
    train_sphere.py:62-80   deform_meshes, sample_points
    train_gcn.py:56-95      sample_predict_points, get_vp_meshes, compose_vp_meshes, calculate_emd_loss
Their text is read from /root/reference at build time, compiled (unmodified) with compile(), and only the marshalled
code objects are written - like oracle/_ref/libemd_ref.so, a binary that travels to the GPU box while no reference
source enters the repository.  tests/test_gpu_train_loop_integration.py executes them against the drop-in `modules`
package, which is what INTEGRATION.md's "stays as written" column claims.  No-op where the reference is absent.
"""
import marshal
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("VPN_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref", "train_helpers.bin")
CHUNKS = {"train": ("train.py", 105, 195), "train_sphere": ("train_sphere.py", 62, 80), "train_gcn": ("train_gcn.py", 56, 95)}


def main():
    if not os.path.isdir(REF):
        print(f"build_ref_train_helpers: {REF} not present - keeping any prebuilt {OUT}")
        return 0
    blob = {"python": sys.version_info[:2]}
    for key, (rel, lo, hi) in CHUNKS.items():
        with open(os.path.join(REF, rel)) as fh:
            lines = fh.readlines()[lo - 1:hi]
        code = compile("".join(lines), f"<reference {rel}:{lo}-{hi}>", "exec")
        blob[key] = marshal.dumps(code)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "wb") as fh:
        marshal.dump(blob, fh)
    print(f"build_ref_train_helpers: wrote {OUT} ({os.path.getsize(OUT)} bytes)")
    return 0


def load(key):
    """Code object of one chunk (for exec in a namespace that provides the names the script imports)."""
    with open(OUT, "rb") as fh:
        blob = marshal.load(fh)
    if tuple(blob["python"]) != tuple(sys.version_info[:2]):
        raise RuntimeError(f"train_helpers.bin was built by Python {blob['python']}, this is {sys.version_info[:2]}")
    return marshal.loads(blob[key])


if __name__ == "__main__":
    sys.exit(main())
