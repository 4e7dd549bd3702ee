"""TEST INFRASTRUCTURE: a reader for the subset of the .pbrt format that rs_pbrt_b200/pbrt_export.py writes, rebuilding a HostScene
from the file -- so that tests can check that an exported scene still IS the scene (export -> read -> same description, same render).
Not a general pbrt parser (the product has none: parsing stays in rs_pbrt)."""
import re
import struct
import zlib
from pathlib import Path

import numpy as np

from rs_pbrt_b200 import HostScene

TOK = re.compile(r'"[^"]*"|\[|\]|[^\s\[\]"]+')


def read_png(path):
    b = Path(path).read_bytes()
    assert b[:8] == b"\x89PNG\r\n\x1a\n"
    i, w, h, data = 8, 0, 0, b""
    while i < len(b):
        n, tag = struct.unpack(">I4s", b[i:i + 8])
        body = b[i + 8:i + 8 + n]
        if tag == b"IHDR":
            w, h = struct.unpack(">II", body[:8])
        elif tag == b"IDAT":
            data += body
        i += 12 + n
    raw = zlib.decompress(data)
    rows = [np.frombuffer(raw[y * (1 + 3 * w) + 1:(y + 1) * (1 + 3 * w)], np.uint8) for y in range(h)]
    return np.stack(rows).reshape(h, w, 3)


def read_ply(path):
    b = Path(path).read_bytes()
    head, body = b.split(b"end_header\n", 1)
    lines = head.decode().splitlines()
    nv = int([l for l in lines if l.startswith("element vertex")][0].split()[-1])
    nf = int([l for l in lines if l.startswith("element face")][0].split()[-1])
    props = [l.split()[-1] for l in lines if l.startswith("property float")]
    v = np.frombuffer(body[:nv * 4 * len(props)], "<f4").reshape(nv, len(props))
    faces = np.frombuffer(body[nv * 4 * len(props):], dtype=[("n", "u1"), ("v", "<i4", 3)], count=nf)
    P = v[:, :3]
    N = v[:, props.index("nx"):props.index("nx") + 3] if "nx" in props else None
    UV = v[:, props.index("u"):props.index("u") + 2] if "u" in props else None
    return faces["v"].reshape(-1).astype(np.uint32), P, N, UV


def _params(tokens):
    """["type name", value or [values], ...] -> {name: (type, list)}"""
    out, i = {}, 0
    while i < len(tokens):
        ty, name = tokens[i].strip('"').split()
        i += 1
        if tokens[i] == "[":
            j = tokens.index("]", i)
            vals = tokens[i + 1:j]
            i = j + 1
        else:
            vals = [tokens[i]]
            i += 1
        out[name] = (ty, [v.strip('"') if v.startswith('"') else float(v) for v in vals])
    return out


MAT = {"matte": 0, "plastic": 1, "metal": 2, "mirror": 3, "glass": 4, "uber": 5, "substrate": 6, "translucent": 7}
LAYOUT = {0: ["Kd", "sigma"], 1: ["Kd", "Ks", "roughness"], 2: ["eta", "k", "uroughness", "vroughness"], 3: ["Kr"], 4: ["Kr", "Kt", "index", "uroughness", "vroughness"],
          5: ["Kd", "Ks", "Kr", "Kt", "opacity", "uroughness", "vroughness", "index"], 6: ["Kd", "Ks", "uroughness", "vroughness"],
          7: ["Kd", "Ks", "reflect", "transmit", "roughness"]}
OFF = {0: [0, 3], 1: [0, 3, 6], 2: [0, 3, 6, 7], 3: [0], 4: [0, 3, 6, 7, 8], 5: [0, 3, 6, 9, 12, 15, 16, 17], 6: [0, 3, 6, 7], 7: [0, 3, 6, 9, 12]}
REMAP = {1: 7, 2: 8, 4: 9, 5: 18, 6: 8, 7: 13}
WRAP = {"repeat": 0, "black": 1, "clamp": 2}


def read(path, n_threads=4):
    path = Path(path)
    toks = TOK.findall("\n".join(l.split("#")[0] for l in path.read_text().splitlines()))
    # split into directives: a directive starts with a bare capitalised word
    dirs, cur = [], None
    depth = 0
    for t in toks:
        if depth == 0 and re.fullmatch(r"[A-Z][A-Za-z]+", t):
            cur = [t]
            dirs.append(cur)
        else:
            cur.append(t)
            depth += (t == "[") - (t == "]")
    h = HostScene()
    tex = {}
    state = [dict(material=None, emit=None, two_sided=False, reverse=False, samples=1, ctm=None, named={})]
    mat_cache = {}
    pending = dict(look=None, cam=None, sampler=None, integ=None, filt=("box", 0.5, 0.5, 2.0), film=None)
    objects = {}

    def flush_pre():
        f = pending["film"]
        h.look_at(*pending["look"])
        h.film(f["x"], f["y"], crop=f["crop"], filter=pending["filt"][0], xwidth=pending["filt"][1], ywidth=pending["filt"][2], alpha=pending["filt"][3],
               max_sample_luminance=f["msl"])
        c = pending["cam"]
        h.camera(fov=c.get("fov", 90.0), lensradius=c.get("lensradius", 0.0), focaldistance=c.get("focaldistance", 1e6), screenwindow=c.get("screenwindow"))
        s = pending["sampler"]
        h.sampler(s["n"], name=s["name"], samplepixelcenter=s["center"])

    def get_material(decl, named):
        name, ps = decl
        if name == "none":
            return -1
        if name == "mix":  # api.rs:678-705: "namedmaterial1" / "namedmaterial2" are looked up among the named materials of the graphics state
            m1 = get_material(named[ps["namedmaterial1"][1][0]], named)
            m2 = get_material(named[ps["namedmaterial2"][1][0]], named)
            key = repr(("mix", m1, m2, ps.get("amount")))
            if key not in mat_cache:
                mat_cache[key] = h.material_mix(m1, m2, ps["amount"][1] if "amount" in ps else (0.5, 0.5, 0.5))
            return mat_cache[key]
        key = repr(decl)
        if key in mat_cache:
            return mat_cache[key]
        kind = MAT[name]
        params = np.zeros(24, np.float32)
        tb, bump = {}, None
        for g, (pn, off) in enumerate(zip(LAYOUT[kind], OFF[kind])):
            if pn in ps:
                ty, v = ps[pn]
                if ty == "texture":
                    tb[g] = tex[v[0]]
                else:
                    params[off:off + len(v)] = v
        if kind in REMAP:
            params[REMAP[kind]] = 1.0 if ps.get("remaproughness", ("bool", ["true"]))[1][0] == "true" else 0.0
        if "bumpmap" in ps:
            bump = tex[ps["bumpmap"][1][0]]
        n_used = 24
        m = h.material(kind, params[:n_used], textures=tb, bump=bump)
        mat_cache[key] = m
        return m

    for d in dirs:
        k, a = d[0], d[1:]
        st = state[-1]
        if k == "LookAt":
            v = [float(x) for x in a]
            pending["look"] = (v[0:3], v[3:6], v[6:9])
        elif k == "Camera":
            p = _params(a[1:])
            pending["cam"] = {n: (v[1] if n == "screenwindow" else v[1][0]) for n, v in p.items()}
        elif k == "Sampler":
            p = _params(a[1:])
            pending["sampler"] = dict(name=a[0].strip('"'), n=int(p["pixelsamples"][1][0]), center=p.get("samplepixelcenter", ("bool", ["false"]))[1][0] == "true")
        elif k == "Integrator":
            pending["integ"] = (a[0].strip('"'), _params(a[1:]))
        elif k == "PixelFilter":
            p = _params(a[1:])
            pending["filt"] = (a[0].strip('"'), p["xwidth"][1][0], p["ywidth"][1][0], p.get("alpha", ("float", [2.0]))[1][0])
        elif k == "Film":
            p = _params(a[1:])
            pending["film"] = dict(x=int(p["xresolution"][1][0]), y=int(p["yresolution"][1][0]), crop=p["cropwindow"][1] if "cropwindow" in p else None,
                                   msl=p["maxsampleluminance"][1][0] if "maxsampleluminance" in p else float("inf"))
        elif k == "WorldBegin":
            pass
        elif k == "AttributeBegin":
            state.append(dict(st))
        elif k == "AttributeEnd":
            state.pop()
        elif k == "Transform":
            st["ctm"] = np.array([float(x) for x in a[1:-1]], np.float32).reshape(4, 4).T
        elif k == "ReverseOrientation":
            st["reverse"] = True
        elif k == "Material":
            st["material"] = (a[0].strip('"'), _params(a[1:]))
        elif k == "MakeNamedMaterial":
            p = _params(a[1:])
            st["named"] = dict(st["named"])  # (the graphics state is copied by AttributeBegin: do not write into the outer block's table)
            st["named"][a[0].strip('"')] = (p.pop("type")[1][0], p)
        elif k == "AreaLightSource":
            p = _params(a[1:])
            st["emit"] = p["L"][1]
            st["two_sided"] = p.get("twosided", ("bool", ["false"]))[1][0] == "true"
            st["samples"] = int(p.get("samples", ("integer", [1]))[1][0])
        elif k == "Texture":
            name, kind, cls = (x.strip('"') for x in a[:3])
            p = _params(a[3:])
            fv = kind == "float"
            if cls == "imagemap":
                img = read_png(path.parent / p["filename"][1][0]).astype(np.float32) / 255.0
                t = h.texture_image(img, trilinear=p["trilinear"][1][0] == "true", max_anisotropy=p["maxanisotropy"][1][0], wrap=WRAP[p["wrap"][1][0]], scale=p["scale"][1][0],
                                    gamma=p["gamma"][1][0] == "true", uscale=p.get("uscale", (0, [1.0]))[1][0], vscale=p.get("vscale", (0, [1.0]))[1][0],
                                    udelta=p.get("udelta", (0, [0.0]))[1][0], vdelta=p.get("vdelta", (0, [0.0]))[1][0], float_valued=fv)
                if "mapping" in p:
                    mp = p["mapping"][1][0]
                    h.texture_mapping(t, mp, np.array(p["v1"][1] + p["v2"][1], np.float32) if mp == "planar" else st["ctm"])
            elif cls == "constant":
                t = h.texture_constant(p["value"][1] if not fv else [p["value"][1][0]], float_valued=fv)
            elif cls == "scale":
                t = h.texture_scale(tex[p["tex1"][1][0]], tex[p["tex2"][1][0]])
            else:
                t = h.texture_mix(tex[p["tex1"][1][0]], tex[p["tex2"][1][0]], tex[p["amount"][1][0]])
            tex[name] = t
        elif k == "Shape":
            p = _params(a[1:])
            if a[0].strip('"') == "plymesh":
                idx, P, N, UV = read_ply(path.parent / p["filename"][1][0])
                S = None
            else:
                idx = np.array(p["indices"][1], np.uint32)
                P = np.array(p["P"][1], np.float32).reshape(-1, 3)
                N = np.array(p["N"][1], np.float32).reshape(-1, 3) if "N" in p else None
                S = np.array(p["S"][1], np.float32).reshape(-1, 3) if "S" in p else None
                UV = np.array(p["uv"][1], np.float32).reshape(-1, 2) if "uv" in p else None
            h.light_samples(st["samples"])
            m = h.trianglemesh(idx, P, N=N, S=S, UV=UV, material=get_material(st["material"], st["named"]), emit=st["emit"], two_sided=st["two_sided"], reverse_orientation=st["reverse"])
            if "alpha" in p or "shadowalpha" in p:
                h.mesh_alpha(m, alpha=tex[p["alpha"][1][0]] if "alpha" in p else None, shadow_alpha=tex[p["shadowalpha"][1][0]] if "shadowalpha" in p else None)
        elif k == "ObjectBegin":
            objects[a[0].strip('"')] = h.object_begin()
        elif k == "ObjectEnd":
            h.object_end()
        elif k == "ObjectInstance":
            h.object_instance(objects[a[0].strip('"')], st["ctm"])
        elif k == "LightSource":
            kind = a[0].strip('"')
            p = _params(a[1:])
            sc = p["scale"][1] if "scale" in p else None
            if kind == "point":
                h.light_point(p["from"][1], p["I"][1], scale=sc)
            elif kind == "spot":
                h.light_spot(p["from"][1], p["to"][1], p["I"][1], scale=sc, coneangle=p["coneangle"][1][0], conedeltaangle=p["conedeltaangle"][1][0])
            elif kind == "distant":
                h.light_distant(p["from"][1], p["to"][1], p["L"][1], scale=sc)
            else:
                assert "mapname" not in p, "environment maps are not read back (RGBE is lossy)"
                h.light_samples(int(p.get("samples", ("integer", [1]))[1][0]))
                h.light_infinite(p["L"][1], scale=sc, light_to_world=None if st["ctm"] is None else st["ctm"][:3, :3])
        elif k == "WorldEnd":
            flush_pre()
            name, p = pending["integ"]
            pb = [int(x) for x in p["pixelbounds"][1]] if "pixelbounds" in p else None
            if name == "path":
                h.integrator(maxdepth=int(p["maxdepth"][1][0]), rrthreshold=p["rrthreshold"][1][0], lightsamplestrategy=p["lightsamplestrategy"][1][0], pixelbounds=pb)
            elif name == "ao":
                h.integrator_ao(nsamples=int(p["nsamples"][1][0]), cossample=p["cossample"][1][0] == "true")
            elif name == "directlighting":
                h.integrator_direct(maxdepth=int(p["maxdepth"][1][0]), strategy=p["strategy"][1][0], pixelbounds=pb)
            else:
                h.integrator_whitted(maxdepth=int(p["maxdepth"][1][0]), pixelbounds=pb)
            h.world_end(n_threads=n_threads)
    return h
