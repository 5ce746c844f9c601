#!/usr/bin/env python3
"""Derive rust/src/ffi.rs (the `extern "C"` block of the Rust crate) from include/kzg_b200.h.

    python tools/gen_ffi_rs.py            # rewrite rust/src/ffi.rs
    python tools/gen_ffi_rs.py --check    # exit 1 if the committed file differs from what the header gives

rustc is not available where this repo is built, so the binding cannot be compiled here; instead it is derived
mechanically from the header and tests/test_ffi_matches_header.py checks the committed file against the header
(names, arity, C types) on every test run."""
import os
import re
import sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
HEADER = os.path.join(ROOT, "include", "kzg_b200.h")
OUT = os.path.join(ROOT, "rust", "src", "ffi.rs")

# C parameter / return type (after array parameters decay to pointers) -> Rust
TYPES = {
    "int": "c_int", "size_t": "usize", "uint64_t": "u64", "void": "()",
    "const char *": "*const c_char", "void *": "*mut c_void",
    "const uint8_t *": "*const u8", "uint8_t *": "*mut u8",
    "const uint32_t *": "*const u32", "uint32_t *": "*mut u32",
    "const int32_t *": "*const i32", "int32_t *": "*mut i32",
    "uint64_t *": "*mut u64", "double *": "*mut f64", "int *": "*mut c_int",
    "kzg_b200_ctx *": "*mut KzgB200Ctx", "const kzg_b200_ctx *": "*const KzgB200Ctx", "kzg_b200_ctx **": "*mut *mut KzgB200Ctx",
}


def strip_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def parse_header(path=HEADER):
    """-> (prototypes, constants): [(name, ret_ctype, [(param_name, ctype), ...])], [(NAME, value)]"""
    text = strip_comments(open(path).read())
    consts = [(m.group(1), int(m.group(2))) for m in re.finditer(r"#define\s+(KZG_B200_[A-Z0-9_]+)\s+(\d+)\s*$", text, flags=re.M)]
    for body in re.findall(r"enum\s*\{(.*?)\}", text, flags=re.S):
        for m in re.finditer(r"(KZG_B200_[A-Z0-9_]+)\s*=\s*(\d+)", body):
            consts.append((m.group(1), int(m.group(2))))
    protos = []
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ ]*?[\s\*]+)(kzg_b200_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", text):
        ret, name, params = normalise(m.group(1)), m.group(2), m.group(3)
        plist = []
        for p in split_params(params):
            pm = re.match(r"^(.*?)([A-Za-z_][A-Za-z0-9_]*)\s*(\[\d*\])?$", p.strip())
            ctype, pname, arr = pm.group(1), pm.group(2), pm.group(3)
            ctype = normalise(ctype + (" *" if arr else ""))
            plist.append((pname, ctype))
        protos.append((name, ret, plist))
    return protos, consts


def split_params(params):
    params = " ".join(params.split())
    if params in ("", "void"):
        return []
    return [p for p in params.split(",")]


def normalise(ctype):
    ctype = " ".join(ctype.replace("*", " * ").split())
    ctype = re.sub(r"\s*\*\s*\*", " **", ctype)
    ctype = re.sub(r"(?<!\*)\s*\*$", " *", ctype)
    return ctype.replace(" * *", " **").strip()


def rust_type(ctype):
    if ctype not in TYPES:
        raise SystemExit("no Rust mapping for C type %r" % ctype)
    return TYPES[ctype]


def render():
    protos, consts = parse_header()
    out = ["//! `extern \"C\"` declarations of libkzg_b200.so -- GENERATED from include/kzg_b200.h by tools/gen_ffi_rs.py,",
           "//! do not edit (tests/test_ffi_matches_header.py compares this file with the header).",
           "#![allow(dead_code)]",
           "use std::os::raw::{c_char, c_int, c_void};", "",
           "/// Opaque context (`kzg_b200_ctx`): plays the role of the reference's `KzgSettings` and owns all device memory.",
           "#[repr(C)]", "pub struct KzgB200Ctx {", "    _private: [u8; 0],", "}", ""]
    for name, value in consts:
        ty = "usize" if name.startswith(("KZG_B200_BYTES_", "KZG_B200_NUM_")) else "c_int"
        out.append("pub const %s: %s = %d;" % (name, ty, value))
    out += ["", "#[link(name = \"kzg_b200\")]", "extern \"C\" {"]
    for name, ret, params in protos:
        args = ", ".join("%s: %s" % (pn, rust_type(ct)) for pn, ct in params)
        r = rust_type(ret)
        out.append("    pub fn %s(%s)%s;" % (name, args, "" if r == "()" else " -> " + r))
    out += ["}", ""]
    return "\n".join(out)


if __name__ == "__main__":
    text = render()
    if "--check" in sys.argv:
        sys.exit(0 if os.path.exists(OUT) and open(OUT).read() == text else 1)
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as fh:
        fh.write(text)
    print(OUT)
