"""The R drop-in package under rpkg/ must (a) keep everything the reference package exports and imports, so that the
untouched R files (clusterbreak.R, plotting.R, similarity.R) keep working after the files are merged in, and (b) be
self-contained: every path src/Makevars names resolves inside rpkg/, also after the symbolic links are turned into
files the way `R CMD build` does.  CPU only; the reference tree is consulted when it is present (build container)."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RPKG = os.path.join(ROOT, "rpkg")
REF = "/root/reference"

# the reference's exported surface (reference NAMESPACE:3-19), kept here so the check also runs without /root/reference
REF_EXPORTS = {"apply_hash", "clusterbreak", "clusterconsensus", "compute_distance_matrix", "compute_signature_matrix",
               "compute_similarity_stats", "consensusplot", "create_char_matrix", "create_hash_parameters", "create_vocab",
               "louvain_mod", "minhash", "netcluster", "plot_similarity_matrix", "shingle", "similarityMH", "similarityNW"}


def directives(path):
    out = set()
    for line in open(path):
        line = line.split("#")[0].strip()
        if line:
            out.add(re.sub(r"\s+", "", line))
    return out


def test_namespace_is_a_superset_of_the_reference():
    mine = directives(os.path.join(RPKG, "NAMESPACE"))
    exports = {m.group(1) for d in mine for m in [re.match(r"export\((\w+)\)$", d)] if m}
    assert REF_EXPORTS <= exports
    assert {"similarityMH_edges", "similarityNW_edges", "netcluster_edges"} <= exports
    assert "useDynLib(DynaAlign,.registration=TRUE)" in mine
    if os.path.exists(os.path.join(REF, "NAMESPACE")):
        ref = directives(os.path.join(REF, "NAMESPACE"))
        assert len([d for d in ref if d.startswith("export(")]) == 17
        missing = ref - mine
        assert not missing, "rpkg/NAMESPACE dropped: %s" % sorted(missing)


def dcf_fields(path):
    fields, key = {}, None
    for line in open(path):
        if line[:1] in " \t" and key:
            fields[key] += " " + line.strip()
        elif ":" in line:
            key, val = line.split(":", 1)
            fields[key.strip()] = val.strip()
    return fields


def test_description_keeps_the_reference_fields():
    mine = dcf_fields(os.path.join(RPKG, "DESCRIPTION"))
    pkgs = lambda s: {p.strip().split(" ")[0] for p in s.split(",") if p.strip()}
    assert {"Rcpp", "igraph", "Biostrings", "DECIPHER", "stats"} <= pkgs(mine["Imports"])
    assert mine["LinkingTo"] == "Rcpp" and mine["LazyData"] == "true" and "SystemRequirements" in mine
    if os.path.exists(os.path.join(REF, "DESCRIPTION")):
        ref = dcf_fields(os.path.join(REF, "DESCRIPTION"))
        for key in ("Package", "Version", "License", "Encoding", "LazyData", "LinkingTo", "Depends", "VignetteBuilder",
                    "Config/testthat/edition"):
            assert mine.get(key) == ref.get(key), key
        for key in ("Imports", "Suggests"):
            assert pkgs(ref[key]) <= pkgs(mine[key]), key
        assert re.sub(r"\s+", "", mine["Authors@R"]) == re.sub(r"\s+", "", ref["Authors@R"])


def makevars_paths():
    text = open(os.path.join(RPKG, "src", "Makevars")).read()
    assert ".." not in re.sub(r"#.*", "", text), "Makevars must not reach outside the package"
    objs = re.search(r"^DYNA_OBJS\s*=\s*(.*)$", text, re.M).group(1).split()
    hdrs = re.search(r"^DYNA_HDRS\s*=\s*(.*)$", text, re.M).group(1).split()
    return [o[:-2] + ".cu" for o in objs] + hdrs


def test_makevars_is_self_contained(tmp_path):
    paths = makevars_paths()
    assert len(paths) == 16
    for rel in paths:
        assert os.path.exists(os.path.join(RPKG, "src", rel)), rel
    # what `R CMD build` ships: links resolved into files; nothing may point outside the copy afterwards
    dst = str(tmp_path / "DynaAlign")
    shutil.copytree(RPKG, dst, symlinks=False)
    for base, _, files in os.walk(dst):
        for f in files:
            assert not os.path.islink(os.path.join(base, f))
    src = os.path.join(dst, "src")
    for rel in paths:
        full = os.path.join(src, rel)
        assert os.path.isfile(full), rel
        for inc in re.findall(r'#include\s+"([^"]+)"', open(full).read()):
            assert os.path.isfile(os.path.join(os.path.dirname(full), inc)), "%s includes %s" % (rel, inc)
    # the shims find the public header through PKG_CPPFLAGS = -Idyna
    assert '#include "dynaalign_b200.h"' in open(os.path.join(src, "dyna_shims.cpp")).read()
    assert os.path.isfile(os.path.join(src, "dyna", "dynaalign_b200.h"))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="no nvcc")
def test_standalone_copy_compiles_one_translation_unit(tmp_path):
    # compile the smallest CUDA source from the resolved copy with the Makevars flags (nvcc cross-compiles on CPU)
    dst = str(tmp_path / "DynaAlign")
    shutil.copytree(RPKG, dst, symlinks=False)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-ccbin", "/usr/bin/g++", "-Xcompiler",
           "-fPIC", "-c", "dyna/probe.cu", "-o", "dyna/probe.o"]
    subprocess.check_call(cmd, cwd=os.path.join(dst, "src"))
    assert os.path.getsize(os.path.join(dst, "src", "dyna", "probe.o")) > 0
