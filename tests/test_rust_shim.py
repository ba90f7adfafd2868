"""The Rust bindings under rust/ cannot be compiled here (no cargo/rustc): check them against the C header instead.
Every `pub fn asrb_*` in ffi.rs must be declared in include/asr_b200.h with the same number of parameters, every
ASRB_API function must be bound, and the #[repr(C)] struct must list the asrb_dims fields in header order."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _c_decls():
    src = open(os.path.join(ROOT, "include", "asr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    out = {}
    for m in re.finditer(r"ASRB_API\s+[\w\s\*]+?\b(asrb_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
    return src, out


def _rust_decls():
    src = open(os.path.join(ROOT, "rust", "src", "backend", "b200", "ffi.rs")).read()
    out = {}
    for m in re.finditer(r"pub fn (asrb_\w+)\s*\(([^)]*)\)", src, flags=re.S):
        args = m.group(2).strip()
        out[m.group(1)] = 0 if not args else len([a for a in args.split(",") if a.strip()])
    return src, out


def test_every_header_function_is_bound_with_the_same_arity():
    _, c = _c_decls()
    _, r = _rust_decls()
    assert len(c) >= 20
    assert set(c) == set(r), (sorted(set(c) - set(r)), sorted(set(r) - set(c)))
    for name, n in c.items():
        assert r[name] == n, (name, n, r[name])


def test_dims_struct_field_order():
    csrc, _ = _c_decls()
    body = re.search(r"typedef struct asrb_dims \{(.*?)\} asrb_dims;", csrc, flags=re.S).group(1)
    c_fields = [f.strip() for decl in re.findall(r"(?:int32_t|double)\s+([^;]+);", body) for f in decl.split(",")]
    rsrc, _ = _rust_decls()
    rbody = re.search(r"pub struct AsrbDims \{(.*?)\n\}", rsrc, flags=re.S).group(1)
    r_fields = re.findall(r"pub (\w+):", rbody)
    assert c_fields == r_fields
    assert len(c_fields) == 20


def test_status_codes_match():
    csrc, _ = _c_decls()
    rsrc, _ = _rust_decls()
    for name in ("ASRB_OK", "ASRB_ERR_INVALID", "ASRB_ERR_CUDA", "ASRB_ERR_IO", "ASRB_ERR_STATE", "ASRB_DT_F32", "ASRB_DT_BF16", "ASRB_DT_F16"):
        cv = int(re.search(rf"#define {name} (\d+)", csrc).group(1))
        rv = int(re.search(rf"pub const {name}: c_int = (\d+);", rsrc).group(1))
        assert cv == rv, name
