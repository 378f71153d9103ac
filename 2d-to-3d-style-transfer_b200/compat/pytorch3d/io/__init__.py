"""OBJ / MTL / texture I/O on either side of the hot path (first_approach.py:83-88, 225;
second_approach.py:77-82, 202).  Runs once per job on the host; restated from the published format
rules: 0-based int64 indices, negative (relative) indices resolved, polygons
fan-triangulated (v0, v_i, v_i+1), texture images returned as float32 (H,W,3) in [0,1] with the top row
first (the v flip lives in the sampler, not here)."""
from __future__ import annotations

import os
from collections import namedtuple
from typing import Dict

import numpy as np
import torch

Faces = namedtuple("Faces", "verts_idx normals_idx textures_idx materials_idx")
Properties = namedtuple("Properties", "normals verts_uvs material_colors texture_images texture_atlas")


def _resolve(tok: str, count: int) -> int:
    """OBJ index (1-based, or negative = relative to the end) -> 0-based; empty -> -1."""
    if tok == "":
        return -1
    i = int(tok)
    return i - 1 if i > 0 else count + i


def _read_mtl(path: str):
    colors: Dict[str, Dict[str, torch.Tensor]] = {}
    maps: Dict[str, str] = {}
    name = None
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t or t[0].startswith("#"):
                continue
            if t[0] == "newmtl":
                name = " ".join(t[1:])
                colors[name] = {}
            elif name is None:
                continue
            elif t[0] == "map_Kd":
                maps[name] = line.split(None, 1)[1].strip()
            elif t[0] in ("Ka", "Kd", "Ks"):
                key = {"Ka": "ambient_color", "Kd": "diffuse_color", "Ks": "specular_color"}[t[0]]
                colors[name][key] = torch.tensor([float(x) for x in t[1:4]], dtype=torch.float32)
            elif t[0] == "Ns":
                colors[name]["shininess"] = torch.tensor([float(t[1])], dtype=torch.float32)
    return colors, maps


def _read_image(path: str) -> torch.Tensor:
    from PIL import Image
    with Image.open(path) as im:
        arr = np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0
    return torch.from_numpy(arr.copy())


def load_obj(f, load_textures: bool = True, create_texture_atlas: bool = False, device="cpu", **unused):
    """-> (verts (V,3) f32, Faces(verts_idx, normals_idx, textures_idx, materials_idx) int64 (F,3),
    Properties(normals, verts_uvs, material_colors, texture_images, texture_atlas))."""
    if create_texture_atlas:
        raise NotImplementedError("texture atlases are not implemented")
    path = os.fspath(f)
    base = os.path.dirname(path)
    verts, uvs, normals = [], [], []
    f_v, f_t, f_n, f_m = [], [], [], []
    mtl_files, material_names, current = [], [], -1
    with open(path) as fh:
        for line in fh:
            t = line.split()
            if not t:
                continue
            tag = t[0]
            if tag == "v":
                verts.append((float(t[1]), float(t[2]), float(t[3])))
            elif tag == "vt":
                uvs.append((float(t[1]), float(t[2])))
            elif tag == "vn":
                normals.append((float(t[1]), float(t[2]), float(t[3])))
            elif tag == "mtllib":
                mtl_files.append(line.split(None, 1)[1].strip())
            elif tag == "usemtl":
                name = " ".join(t[1:])
                if name not in material_names:
                    material_names.append(name)
                current = material_names.index(name)
            elif tag == "f":
                corners = []
                for c in t[1:]:
                    parts = (c.split("/") + ["", ""])[:3]
                    corners.append((_resolve(parts[0], len(verts)), _resolve(parts[1], len(uvs)),
                                    _resolve(parts[2], len(normals))))
                if len(corners) < 3:
                    raise ValueError(f"face with fewer than 3 vertices: {line.strip()!r}")
                for j in range(1, len(corners) - 1):                      # fan triangulation
                    tri = (corners[0], corners[j], corners[j + 1])
                    f_v.append([c[0] for c in tri])
                    f_t.append([c[1] for c in tri])
                    f_n.append([c[2] for c in tri])
                    f_m.append(current)

    def idx(rows):
        return torch.tensor(rows, dtype=torch.int64, device=device).reshape(-1, 3)

    verts_t = torch.tensor(verts, dtype=torch.float32, device=device).reshape(-1, 3)
    uvs_t = torch.tensor(uvs, dtype=torch.float32, device=device).reshape(-1, 2) if uvs else None
    normals_t = torch.tensor(normals, dtype=torch.float32, device=device).reshape(-1, 3) if normals else None
    material_colors, texture_images = {}, {}
    if load_textures:
        for m in mtl_files:
            mpath = os.path.join(base, m)
            if not os.path.exists(mpath):
                continue
            colors, maps = _read_mtl(mpath)
            for name, c in colors.items():
                material_colors[name] = {k: v.to(device) for k, v in c.items()}
            for name, rel in maps.items():
                ipath = os.path.join(base, rel)
                if os.path.exists(ipath):
                    texture_images[name] = _read_image(ipath).to(device)
    faces = Faces(idx(f_v), idx(f_n), idx(f_t), torch.tensor(f_m, dtype=torch.int64, device=device))
    aux = Properties(normals_t, uvs_t, material_colors or None, texture_images or None, None)
    return verts_t, faces, aux


def save_obj(path, verts, faces, verts_uvs=None, faces_uvs=None, texture_map=None, decimal_places: int = 6):
    """Writes OBJ (+ MTL + PNG when a texture map is given)."""
    path = os.fspath(path)
    stem = os.path.splitext(os.path.basename(path))[0]
    folder = os.path.dirname(path) or "."
    os.makedirs(folder, exist_ok=True)
    verts = verts.detach().cpu().numpy()
    faces = faces.detach().cpu().numpy()
    has_tex = texture_map is not None and verts_uvs is not None and faces_uvs is not None
    fmt = f"%.{decimal_places}f"
    lines = []
    if has_tex:
        from PIL import Image
        img = (texture_map.detach().clamp(0, 1).cpu().numpy() * 255.0).round().astype(np.uint8)
        Image.fromarray(img).save(os.path.join(folder, stem + ".png"))
        with open(os.path.join(folder, stem + ".mtl"), "w") as fh:
            fh.write("newmtl mesh\nKa 1.0 1.0 1.0\nKd 1.0 1.0 1.0\nKs 0.0 0.0 0.0\nNs 10.0\n" f"map_Kd {stem}.png\n")
        lines += [f"mtllib {stem}.mtl", "usemtl mesh"]
    lines += ["v " + " ".join(fmt % c for c in v) for v in verts]
    if has_tex:
        uv = verts_uvs.detach().cpu().numpy()
        fuv = faces_uvs.detach().cpu().numpy()
        lines += ["vt " + " ".join(fmt % c for c in t) for t in uv]
        lines += ["f " + " ".join(f"{a + 1}/{b + 1}" for a, b in zip(fv, ft)) for fv, ft in zip(faces, fuv)]
    else:
        lines += ["f " + " ".join(str(a + 1) for a in fv) for fv in faces]
    with open(path, "w") as fh:
        fh.write("\n".join(lines) + "\n")


class IO:
    """IO().save_mesh(mesh, path) / IO().load_mesh(path) for .obj files."""

    def save_mesh(self, data, path, binary=None, include_textures: bool = True, **unused):
        if not str(path).lower().endswith(".obj"):
            raise NotImplementedError("only .obj meshes are implemented")
        if len(data) != 1:
            raise NotImplementedError("one mesh per file")
        tex = data.textures
        kw = {}
        if include_textures and tex is not None and hasattr(tex, "maps_padded"):
            kw = dict(verts_uvs=tex.verts_uvs_padded()[0], faces_uvs=tex.faces_uvs_padded()[0],
                      texture_map=tex.maps_padded()[0])
        save_obj(path, data.verts_packed(), data.faces_packed(), **kw)

    def load_mesh(self, path, include_textures: bool = True, device="cpu", **unused):
        from ..renderer import TexturesUV
        from ..structures import Meshes
        verts, faces, aux = load_obj(path, load_textures=include_textures, device=device)
        tex = None
        if include_textures and aux.texture_images and aux.verts_uvs is not None:
            image = list(aux.texture_images.values())[0]
            tex = TexturesUV(maps=image[None], faces_uvs=faces.textures_idx[None], verts_uvs=aux.verts_uvs[None])
        return Meshes(verts=[verts], faces=[faces.verts_idx], textures=tex)


__all__ = ["load_obj", "save_obj", "IO", "Faces", "Properties"]
