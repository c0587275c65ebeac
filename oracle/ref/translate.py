#!/usr/bin/env python3
"""Builds oracle/_ref/libtess_ref_<variant>.so : the REFERENCE'S OWN compute shaders compiled for the host.

TEST INFRASTRUCTURE.  Reads the GLSL sources where they lie under /root/reference/shaders (never copied into the
repository: everything this script writes goes to oracle/_ref/, which is git-ignored), runs them through the C
preprocessor with the macro set src/renderer_raytrace_clusters_tess.cpp:116-141 passes to shaderc, rewrites the few
constructs C++ cannot parse (interface blocks, buffer references, parameter qualifiers, swizzles, float literals, call
sites of subgroup operations) and compiles the result together with oracle/ref/glsl_shim.hpp (GLSL run-time + SIMT
emulator) and oracle/ref/ref_harness.cpp (the dispatch schedule of RendererRayTraceClustersTess::render).

usage: translate.py [--flags N] [--vis-bits B --split-bits B --part-bits B --vert-bits B --megs M] [--textures 0|1]
                    [--cluster-verts N --cluster-tris N] [--split-factor N] [--out PATH]
"""
from __future__ import annotations

import argparse
from concurrent.futures import ThreadPoolExecutor
import hashlib
import os
import re
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SHADERS = os.environ.get("TC_REFERENCE_SHADERS", "/root/reference/shaders")
OUT_DIR = os.path.join(os.path.dirname(HERE), "_ref")

SHADERS = ["instances_classify", "clusters_cull", "build_setup", "cluster_classify", "triangle_split",
           "triangle_tess_template_instantiate", "blas_setup_insertion", "blas_clusters_insert"]

FLAG_PN, FLAG_1X, FLAG_2X, FLAG_CULL, FLAG_ANIM = 1, 2, 4, 8, 16

COLLECTIVES = ["subgroupBallot", "subgroupAny", "subgroupAll", "subgroupElect", "subgroupBroadcastFirst", "subgroupBroadcast", "subgroupShuffle",
               "subgroupAdd", "subgroupInclusiveAdd", "subgroupExclusiveAdd", "subgroupMax", "subgroupMin", "subgroupOr", "subgroupPartitionNV",
               "barrier"]
SWIZZLES = ["xy", "yx", "zw", "xyz", "yzx", "zxy", "xzy", "yxz", "zyx"]


def macro_set(a) -> dict:
    """src/renderer_raytrace_clusters_tess.cpp:116-141"""
    max_vis, max_part = 1 << a.vis_bits, 1 << a.part_bits
    return {
        "CLUSTER_VERTEX_COUNT": a.cluster_verts, "CLUSTER_TRIANGLE_COUNT": a.cluster_tris, "TESSTABLE_SIZE": 11, "TESSTABLE_LOOKUP_SIZE": 16,
        "TARGETS_RASTERIZATION": 0, "TESS_USE_PN": int(bool(a.flags & FLAG_PN)), "TESS_USE_1X_TRANSIENTBUILDS": int(bool(a.flags & FLAG_1X)),
        "TESS_USE_2X_TRANSIENTBUILDS": int(bool(a.flags & FLAG_2X)), "TESS_USE_PERSISTENT_KERNEL": 0,
        "TESS_MAX_SPLIT_FACTOR": max(2, min(a.split_factor, 11)), "TESS_ACTIVE": 1, "MAX_PART_TRIANGLES": max_part, "MAX_VISIBLE_CLUSTERS": max_vis,
        "MAX_SPLIT_TRIANGLES": 1 << a.split_bits, "MAX_GENERATED_CLUSTER_MEGS": a.megs, "MAX_GENERATED_CLUSTERS": max_vis + max_part,
        "MAX_GENERATED_VERTICES": 1 << a.vert_bits, "HAS_DISPLACEMENT_TEXTURES": a.textures, "DO_CULLING": int(bool(a.flags & FLAG_CULL)),
        "DO_ANIMATION": int(bool(a.flags & FLAG_ANIM)), "DEBUG_VISUALIZATION": 0,
    }


def variant_name(macros: dict) -> str:
    key = ";".join(f"{k}={v}" for k, v in sorted(macros.items())) + repr(HIZ_PROGRAMS)
    for f in ("glsl_shim.hpp", "ref_harness.cpp", "translate.py"):
        with open(os.path.join(HERE, f), "rb") as fh:
            key += hashlib.sha1(fh.read()).hexdigest()
    return hashlib.sha1(key.encode()).hexdigest()[:12]


def stage_sources(stage: str):
    """directive-stripped copies (gcc -E rejects #version / #extension) + a stub for the one nvpro_core2 include"""
    os.makedirs(os.path.join(stage, "nvshaders"), exist_ok=True)
    for f in os.listdir(REF_SHADERS):
        if f.endswith((".glsl", ".h")):
            with open(os.path.join(REF_SHADERS, f), encoding="utf-8", errors="replace") as fh:
                txt = fh.read()
            txt = re.sub(r"^[ \t]*#[ \t]*(version|extension)\b.*$", "", txt, flags=re.M)
            with open(os.path.join(stage, f), "w") as fh:
                fh.write(txt)
    with open(os.path.join(stage, "render_shading.glsl"), "w") as fh:
        fh.write("\n")  # shading / AO rays of the closest-hit shader: after the decode the path ends at, not compiled
    with open(os.path.join(stage, "nvshaders", "sky_io.h.slang"), "w") as fh:
        fh.write("struct SkySimpleParameters { vec4 _stub; };\n")  # tail of FrameConstants, never read by the path


# the far-HiZ builder: one shader, two pipelines -- the macro sets of NVHizVK::initPipelines
# (src/nvhiz_vk.cpp:226-240) with the sample's configuration (src/resources.cpp:176-186, nvhiz_vk.cpp:79: 3 levels per
# dispatch, no MSAA, no reversed Z, far pyramid only, mono)
HIZ_COMMON = {"NV_HIZ_LEVELS": 3, "NV_HIZ_MSAA_SAMPLES": 0, "NV_HIZ_REVERSED_Z": 0, "NV_HIZ_NEAR_LEVEL": 0, "NV_HIZ_FAR_LEVEL": 0,
              "NV_HIZ_OUTPUT_NEAR": 0, "NV_HIZ_USE_STEREO": 0}
HIZ_PROGRAMS = [("hiz_first", "nvhiz-update", dict(HIZ_COMMON, NV_HIZ_IS_FIRST=1)), ("hiz_rest", "nvhiz-update", dict(HIZ_COMMON, NV_HIZ_IS_FIRST=0))]


def preprocess(stage: str, shader: str, macros: dict) -> str:
    cmd = ["gcc", "-E", "-P", "-undef", "-nostdinc", "-x", "c", "-I", stage] + [f"-D{k}={v}" for k, v in macros.items()] + [os.path.join(stage, shader if shader.endswith(".glsl") else shader + ".comp.glsl")]
    return subprocess.run(cmd, check=True, capture_output=True, text=True).stdout


def _split_members(body: str):
    return [m.strip() for m in body.split(";") if m.strip()]


RCHIT_CUT = "vec3 oPos = baryWeight.x * gl_HitTriangleVertexPositionsEXT[0]"
RCHIT_EXPORT = """
  // (translate.py) main() is cut where the hit decode ends and shading begins: hand the decoded values to the harness
  _hit_out->mode = mode; _hit_out->clusterID = clusterID; _hit_out->triangleID = triangleID; _hit_out->subTriangleID = subTriangleID;
  _hit_out->cfg = cfg; _hit_out->baseIndices[0] = baseIndices.x; _hit_out->baseIndices[1] = baseIndices.y; _hit_out->baseIndices[2] = baseIndices.z;
  _hit_out->partID = partID; _hit_out->bary[0] = baryWeightBase.x; _hit_out->bary[1] = baryWeightBase.y; _hit_out->bary[2] = baryWeightBase.z;
}
"""


def rchit_prepare(src: str) -> str:
    """closest-hit shader -> a function: ray-tracing built-ins become plain variables the harness sets per hit"""
    src = re.sub(r"spirv_decorate\s*\(.*?\)\s*in\s+int\s+gl_ClusterIDNV_\s*;", "", src, flags=re.S)
    src = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+accelerationStructureEXT\s+\w+\s*;", "", src)
    src = re.sub(r"layout\s*\([^)]*\)\s*rayPayload(In)?EXT\s+\w+\s+\w+\s*;", "", src)
    src = re.sub(r"hitAttributeEXT\s+vec2\s+barycentrics\s*;", "", src)
    cut = src.index(RCHIT_CUT)
    head = ("struct HitOut { uint mode, clusterID, triangleID, subTriangleID, cfg, baseIndices[3], partID; float bary[3]; };\n"
            "static HitOut* _hit_out; static int gl_ClusterIDNV_, gl_PrimitiveID, gl_InstanceID; static vec2 barycentrics;\n")
    return head + src[:cut] + RCHIT_EXPORT + "layout(local_size_x=1) in;\n"


def raster_macro_set(macros: dict) -> dict:
    """src/renderer_raster_clusters_tess.cpp:93-113: the rasteriser's configuration of the same frame (no transient builds,
    batched meshlets on); none of the differences touches a struct layout the harness's buffers depend on"""
    m = dict(macros, TARGETS_RASTERIZATION=1, TESS_RASTER_USE_BATCH=1, TESS_USE_1X_TRANSIENTBUILDS=0, TESS_USE_2X_TRANSIENTBUILDS=0, MESHSHADER_WORKGROUP_SIZE=32)
    for k in ("MAX_GENERATED_CLUSTER_MEGS", "MAX_GENERATED_CLUSTERS", "MAX_GENERATED_VERTICES"):
        m.pop(k, None)
    return m


def raster_task_prepare(src: str) -> str:
    """task shader -> compute-style function: the taskNV output block becomes a plain struct the runner copies out after
    every workgroup, gl_TaskCountNV a variable"""
    src, n = re.subn(r"\bout\s+taskNV\s+TaskExchange\s*\{([^}]*)\}\s*TASK\s*;",
                     r"struct TaskExchange_t {\1};\nstatic TaskExchange_t TASK; static uint gl_TaskCountNV;", src)
    assert n == 1
    return src


def raster_mesh_prepare(src: str) -> str:
    """mesh shader -> compute-style function: the taskNV input block, the per-vertex interface block and the mesh built-ins
    become plain statics the runner fills / copies out around every workgroup"""
    src, n = re.subn(r"\btaskNV\s+in\s+TaskExchange\s*\{([^}]*)\}\s*TASK\s*;", r"struct TaskExchange_t {\1};\nstatic TaskExchange_t TASK;", src)
    assert n == 1
    src, n = re.subn(r"layout\s*\(\s*location\s*=\s*0\s*\)\s*out\s+Interpolants\s*\{([^}]*)\}\s*OUT\s*\[\s*\]\s*;",
                     lambda m: "struct Interpolants_t {" + re.sub(r"\bflat\s+", "", m.group(1)) + "};\nstatic Interpolants_t OUT[96];", src)
    assert n == 1
    src, n = re.subn(r"layout\s*\(\s*max_vertices.*?\)\s*out\s*;", "", src)
    assert n == 1
    src, n = re.subn(r"layout\s*\(\s*triangles\s*\)\s*out\s*;", "", src)
    assert n == 1
    head = ("struct MeshVertex_t { vec4 gl_Position; };\nstatic MeshVertex_t gl_MeshVerticesNV[96];\n"
            "struct MeshPrimitive_t { int gl_PrimitiveID; };\nstatic MeshPrimitive_t gl_MeshPrimitivesNV[128];\n"
            "static uint gl_PrimitiveIndicesNV[3 * 128];\nstatic uint gl_PrimitiveCountNV;\n")
    return head + src


RASTER_MESH_RUNNER = """struct MeshOut { uint primitiveCount; uint indices[3 * 128]; int primitiveIDs[128]; float wPos[96 * 3]; float clip[96 * 4]; uint clusterID[96]; uint instanceID[96]; };
void run_groups_raster_mesh(const void* task, uint numWorkgroups, void* out)  // task: one 200-byte record; out: numWorkgroups MeshOut
{
  static_assert(sizeof(TaskExchange_t) == 196, "TaskExchange");
  memcpy(&TASK, task, 196);
  MeshOut* o = static_cast<MeshOut*>(out);
  for(uint gx = 0; gx < numWorkgroups; gx++)
  {
    memset(gl_MeshVerticesNV, 0, sizeof(gl_MeshVerticesNV)); memset(gl_MeshPrimitivesNV, 0, sizeof(gl_MeshPrimitivesNV));
    memset(gl_PrimitiveIndicesNV, 0, sizeof(gl_PrimitiveIndicesNV)); memset(OUT, 0, sizeof(OUT));
    gl_PrimitiveCountNV = 0;
    Simt::get().runWorkgroup(&shader_main, _local_size_x * _local_size_y, gx, 32, _local_size_x, 0);
    o[gx].primitiveCount = gl_PrimitiveCountNV;
    memcpy(o[gx].indices, gl_PrimitiveIndicesNV, sizeof(gl_PrimitiveIndicesNV));
    for(int t = 0; t < 128; t++) o[gx].primitiveIDs[t] = gl_MeshPrimitivesNV[t].gl_PrimitiveID;
    for(int v = 0; v < 96; v++)
    {
      o[gx].wPos[v * 3 + 0] = OUT[v].wPos.x; o[gx].wPos[v * 3 + 1] = OUT[v].wPos.y; o[gx].wPos[v * 3 + 2] = OUT[v].wPos.z;
      o[gx].clip[v * 4 + 0] = gl_MeshVerticesNV[v].gl_Position.x; o[gx].clip[v * 4 + 1] = gl_MeshVerticesNV[v].gl_Position.y;
      o[gx].clip[v * 4 + 2] = gl_MeshVerticesNV[v].gl_Position.z; o[gx].clip[v * 4 + 3] = gl_MeshVerticesNV[v].gl_Position.w;
      o[gx].clusterID[v] = OUT[v].clusterID; o[gx].instanceID[v] = OUT[v].instanceID;
    }
  }
}"""


RASTER_TASK_RUNNER = """void run_groups_raster_task(uint groupsX, void* out)  // out: 200-byte records (TaskExchange + gl_TaskCountNV)
{
  static_assert(sizeof(TaskExchange_t) == 196, "TaskExchange");
  for(uint gx = 0; gx < groupsX; gx++)
  {
    memset(&TASK, 0, sizeof(TASK));
    gl_TaskCountNV = 0;
    Simt::get().runWorkgroup(&shader_main, _local_size_x * _local_size_y, gx, 32, _local_size_x, 0);
    memcpy(static_cast<char*>(out) + size_t(gx) * 200, &TASK, 196);
    memcpy(static_cast<char*>(out) + size_t(gx) * 200 + 196, &gl_TaskCountNV, 4);
  }
}"""


def to_cpp(src: str, shader: str) -> str:
    out_pre = []  # emitted before the translated text
    if shader == "rchit":
        src = rchit_prepare(src)
    if shader == "raster_task":
        src = raster_task_prepare(src)
    if shader == "raster_mesh":
        src = raster_mesh_prepare(src)

    # ---- float literals: GLSL literals are float, C++ literals are double
    def lit(m):
        s = m.group(0)
        return s if s.endswith(("f", "F")) else s + "f"
    src = re.sub(r"(?<![\w.])(\d+\.\d*|\.\d+)([eE][-+]?\d+)?[fF]?(?![\w.])", lit, src)
    src = re.sub(r"(?<![\w.])(\d+)([eE][-+]?\d+)(?![\w.])", lambda m: m.group(0) + "f", src)

    # ---- buffer references -> typed host pointers
    def bufref(m):
        name, typ = m.group(1), m.group(2)
        return (f"struct {name} {{ {typ}* d; {name}() = default; explicit {name}(uint64_t a) : d(reinterpret_cast<{typ}*>(a)) {{}} "
                f"template <class O, class = decltype(O().d)> explicit {name}(const O& o) : d(reinterpret_cast<{typ}*>(o.d)) {{}} "
                f"explicit operator uint64_t() const {{ return reinterpret_cast<uint64_t>(d); }} }};")
    src = re.sub(r"layout\s*\(\s*buffer_reference[^)]*\)\s*[\w\s]*?\bbuffer\s+(\w+)\s*\{\s*(\w+)\s+d\s*\[\s*\]\s*;\s*\}\s*;", bufref, src)

    # ---- push constants, uniform / storage blocks, samplers, workgroup size
    binds = []  # (type, name, is_array)

    def block(m):
        layout, body, inst = m.group(1), m.group(3), m.group(4)
        members = _split_members(body)
        if "push_constant" in layout:
            fields = " ".join(x + ";" for x in members)
            if not inst:  # members are globals: one struct behind the name `push`, one macro per member
                for mem in members:
                    out_pre.append(f"#define {mem.split()[-1]} (_p_push->m_{mem.split()[-1]})")
                fields = " ".join(" ".join(x.split()[:-1]) + " m_" + x.split()[-1] + ";" for x in members)
                binds.append(("push_t", "push", None))
                return f"struct push_t {{ {fields} }};\nstatic push_t* _p_push;\n"
            binds.append((f"{inst}_t", inst, False))
            return f"struct {inst}_t {{ {fields} }};\nstatic {inst}_t* _p_{inst};\n"
        txt = ""
        for mem in members:
            mm = re.match(r"(?:(?:volatile|coherent|readonly|restrict)\s+)*(\w+)\s+(\w+)\s*(\[\s*\])?$", mem)
            typ, name, arr = mm.group(1), mm.group(2), mm.group(3)
            binds.append((typ, name, bool(arr)))
            txt += f"static {typ}* _p_{name};\n"
        return txt
    src = re.sub(r"layout\s*\(([^)]*)\)\s*(uniform|buffer|readonly\s+buffer|coherent\s+buffer)\s+\w+\s*\{([^}]*)\}\s*(\w*)\s*;", block, src)

    def sampler(m):
        name, arr = m.group(1), m.group(2)
        binds.append(("sampler2D", name, bool(arr)))
        return f"static sampler2D* _p_{name};\n"
    src = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+sampler2D\s+(\w+)\s*(\[\s*\])?\s*;", sampler, src)

    def image(m):
        name, arr = m.group(1), m.group(2)
        binds.append(("image2D", name, bool(arr)))
        return f"static image2D* _p_{name};\n"
    src = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+image2D\s+(\w+)\s*(\[\s*\w*\s*\])?\s*;", image, src)

    local = re.search(r"layout\s*\(\s*local_size_x\s*=\s*(\d+)\s*(?:,\s*local_size_y\s*=\s*(\d+)\s*)?\)\s*in\s*;", src)
    src = src.replace(local.group(0), f"static const uint _local_size_x = {local.group(1)}, _local_size_y = {local.group(2) or 1};")

    for typ, name, arr in binds:
        if arr is not None:
            out_pre.append(f"#define {name} {'_p_' + name if arr else '(*_p_' + name + ')'}")

    # ---- qualifiers
    src = re.sub(r"\bshared\s+", "static ", src)
    src = re.sub(r"\bprecise\s+", "", src)
    src = re.sub(r"\b(inout|out)\s+(\w+)\s+(\w+)\s*\[\s*(\w+)\s*\]", r"\2 (&\3)[\4]", src)
    src = re.sub(r"\b(inout|out)\s+(\w+)\s+(\w+)", r"\2& \3", src)
    src = re.sub(r"([(,]\s*)in\s+(\w+\s+\w+)", r"\1\2", src)
    src = re.sub(r"\[\[\s*(unroll|branch|flatten|loop|dont_unroll)\s*\]\]", "", src)

    # ---- struct constructors (GLSL gives every struct a member-wise constructor)
    def struct(m):
        name, body = m.group(1), m.group(2)
        members = _split_members(body)
        if any("[" in x for x in members) or name.endswith("_t"):
            return m.group(0)
        params, inits = [], []
        for mem in members:
            mm = re.match(r"([\w]+)\s+(\w+)$", mem)
            if not mm:
                return m.group(0)
            params.append(f"const {mm.group(1)}& {mm.group(2)}_")
            inits.append(f"{mm.group(2)}({mm.group(2)}_)")
        return f"struct {name}\n{{{body}  {name}() = default;\n  {name}({', '.join(params)}) : {', '.join(inits)} {{}}\n}}"
    src = re.sub(r"\bstruct\s+(\w+)\s*\{([^{}]*)\}", struct, src)

    # ---- `T x = x[i];` : in GLSL the new name is not yet in scope inside its own initialiser, in C++ it is
    while True:
        m = re.search(r"\b(\w+)\s+(\w+)\s*=\s*\2\s*\[[^\]]*\]\s*;", src)
        if not m:
            break
        name, depth, end = m.group(2), 0, m.end()
        while end < len(src) and depth >= 0:
            depth += {"{": 1, "}": -1}.get(src[end], 0)
            end += 1
        head = f"{m.group(1)} {name}_inner = {src[m.start(2) + len(name):m.end()].split('=', 1)[1]}"
        src = src[:m.start()] + head + re.sub(rf"\b{name}\b", name + "_inner", src[m.end():end]) + src[end:]

    # ---- swizzles -> accessor calls
    src = re.sub(r"\.(" + "|".join(SWIZZLES) + r")\b(?!\s*\()", r".\1()", src)

    # ---- collectives get their call-site id (text order)
    counter = [0]

    def site(m):
        counter[0] += 1
        name, rest = m.group(1), m.group(2)
        return f"{name}_({counter[0]}{'' if rest.strip().startswith(')') else ', '}{rest}"
    src = re.sub(r"\b(" + "|".join(COLLECTIVES) + r")\s*\((\s*\)?)", site, src)

    src = re.sub(r"\bvoid\s+main\s*\(\s*\)", "static void shader_main()", src)

    bind_fn = "\n".join(
        f"  if(!strcmp(name, \"{name}\")) {{ _p_{name} = reinterpret_cast<decltype(_p_{name})>(ptr); return 1; }}" for _, name, _ in binds)
    return f"""// GENERATED by oracle/ref/translate.py from {REF_SHADERS}/{shader}.comp.glsl -- build artefact, do not commit
#include "glsl_shim.hpp"
{chr(10).join(out_pre)}
namespace glsl {{ namespace {{
{src}
}}
int bind_{shader}(const char* name, void* ptr)
{{
{bind_fn}
  return 0;
}}
{RCHIT_SETTER if shader == "rchit" else ""}
{RASTER_TASK_RUNNER if shader == "raster_task" else ""}
{RASTER_MESH_RUNNER if shader == "raster_mesh" else ""}
uint local_size_{shader}() {{ return _local_size_x; }}
void run_{shader}(uint groupsX, uint groupsY)
{{
  for(uint gy = 0; gy < groupsY; gy++)
    for(uint gx = 0; gx < groupsX; gx++)
      Simt::get().runWorkgroup(&shader_main, _local_size_x * _local_size_y, gx, 32, _local_size_x, gy);
}}
}}
"""


RCHIT_SETTER = """void set_hit_rchit(uint clusterID, uint primitiveID, uint instanceID, float b0, float b1, void* out)
{
  gl_ClusterIDNV_ = int(clusterID); gl_PrimitiveID = int(primitiveID); gl_InstanceID = int(instanceID);
  barycentrics = vec2(b0, b1); _hit_out = reinterpret_cast<HitOut*>(out);
}"""


def build(a) -> str:
    macros = macro_set(a)
    name = variant_name(macros)
    out = a.out or os.path.join(OUT_DIR, f"libtess_ref_{name}.so")
    if os.path.exists(out) and not a.force:
        return out
    if not os.path.isdir(REF_SHADERS):
        raise SystemExit(f"{REF_SHADERS} is not present and {out} was not prebuilt")
    gen = os.path.join(OUT_DIR, f"gen_{name}_{os.getpid()}")
    stage = os.path.join(gen, "src")
    shutil.rmtree(gen, ignore_errors=True)
    stage_sources(stage)
    objs = []
    cxx = ["g++", "-std=c++17", "-O1", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-msse4.1", "-w", "-fvisibility=hidden", "-I", HERE,
           "-I", os.path.join(os.path.dirname(os.path.dirname(HERE)), "include")]
    def compile_one(prog):
        sh, file, defs = prog
        cpp = os.path.join(gen, sh + ".cpp")
        with open(cpp, "w") as fh:
            fh.write(to_cpp(preprocess(stage, file, defs), sh))
        obj = cpp[:-4] + ".o"
        r = subprocess.run(cxx + ["-c", cpp, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stderr[:6000])
            raise SystemExit(f"compiling the translated {sh} failed")
        return obj
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        rchit = ("rchit", "render_raytrace_clusters.rchit.glsl", dict(macros, RAYTRACING_PAYLOAD_INDEX=0))
        raster_task = ("raster_task", "render_raster_clusters_batched.task.glsl", raster_macro_set(macros))
        raster_mesh = ("raster_mesh", "render_raster_clusters_batched.mesh.glsl", raster_macro_set(macros))
        objs = list(pool.map(compile_one, [(sh, sh, macros) for sh in SHADERS] + HIZ_PROGRAMS + [rchit, raster_task, raster_mesh]))
    defs = [f"-DREF_{k}={v}" for k, v in macros.items()]
    with open(os.path.join(REF_SHADERS, "shaderio.h")) as fh:  # push-constant ids of build_setup.comp.glsl
        defs += [f"-DREF_{m.group(1)}={m.group(2)}" for m in re.finditer(r"^#define\s+(BUILD_SETUP_\w+)\s+(\d+)", fh.read(), flags=re.M)]
    tmp = out + f".tmp{os.getpid()}"
    r = subprocess.run(cxx + defs + [os.path.join(HERE, "ref_harness.cpp")] + objs + ["-shared", "-o", tmp], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stderr[:6000])
        raise SystemExit("linking the reference-shader library failed")
    os.replace(tmp, out)
    if not a.keep:
        shutil.rmtree(gen, ignore_errors=True)  # only the .so stays (and travels to the GPU box)
    return out


def parser():
    p = argparse.ArgumentParser()
    p.add_argument("--flags", type=int, default=FLAG_PN | FLAG_1X | FLAG_2X)
    p.add_argument("--vis-bits", type=int, default=20)
    p.add_argument("--split-bits", type=int, default=16)
    p.add_argument("--part-bits", type=int, default=20)
    p.add_argument("--vert-bits", type=int, default=24)
    p.add_argument("--megs", type=int, default=1024)
    p.add_argument("--textures", type=int, default=1)
    p.add_argument("--cluster-verts", type=int, default=64)
    p.add_argument("--cluster-tris", type=int, default=64)
    p.add_argument("--split-factor", type=int, default=8)
    p.add_argument("--out", default=None)
    p.add_argument("--force", action="store_true")
    p.add_argument("--keep", action="store_true", help="keep the generated C++ under oracle/_ref/gen_* (debugging)")
    return p


if __name__ == "__main__":
    print(build(parser().parse_args()))
