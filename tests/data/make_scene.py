"""Writes the small kazen XML test scene (tests/data/box/) used by the host tests: OBJ meshes with
quads, shared vertices, normals and uvs, a PNG texture, every transform op, a nested normal map."""
import os
import struct
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def write_png(path, img8):
    h, w, _ = img8.shape
    raw = b"".join(b"\x00" + img8[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xFFFFFFFF)
    open(path, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) + chunk(b"IDAT", zlib.compress(raw)) + chunk(b"IEND", b""))


def sphere_obj(path, nu=16, nv=8):
    lines = []
    for j in range(nv + 1):
        th = np.pi * j / nv
        for i in range(nu + 1):
            ph = 2 * np.pi * i / nu
            n = (np.sin(th) * np.cos(ph), np.cos(th), np.sin(th) * np.sin(ph))
            lines.append("v %.6f %.6f %.6f" % n)
            lines.append("vn %.6f %.6f %.6f" % n)
            lines.append("vt %.6f %.6f" % (i / nu, 1 - j / nv))
    for j in range(nv):
        for i in range(nu):
            a = j * (nu + 1) + i + 1; b = a + 1; c = a + nu + 1; d = c + 1
            if j == 0:
                lines.append(f"f {b}/{b}/{b} {d}/{d}/{d} {c}/{c}/{c}")
            elif j == nv - 1:
                lines.append(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}")
            else:
                lines.append(f"f {a}/{a}/{a} {b}/{b}/{b} {d}/{d}/{d} {c}/{c}/{c}")      # quad
    open(path, "w").write("\n".join(lines) + "\n")


ROOM = """# open box, quads, no normals / uvs
v -1 -1 -1
v 1 -1 -1
v 1 -1 1
v -1 -1 1
v -1 1 -1
v 1 1 -1
v 1 1 1
v -1 1 1
f 4 3 2 1
f 6 7 8 5
f 8 7 3 4
f 5 8 4 1
f 3 7 6 2
"""
LIGHT = """v -0.3 0.98 -0.3
v 0.3 0.98 -0.3
v 0.3 0.98 0.3
v -0.3 0.98 0.3
vn 0 -1 0
f 1//1 2//1 3//1 4//1
"""
XML = """<?xml version="1.0" ?>
<!-- kazen-b200 host test scene -->
<scene>
	<integrator type="path_mis">
		<integer name="maxDepth" value="4"/>
		<boolean name="regularization" value="true"/>
	</integrator>
	<sampler type="stratified">
		<integer name="sampleCount" value="10"/>
	</sampler>
	<camera type="thinlens">
		<float name="fov" value="39"/>
		<float name="nearClip" value="0.1"/>
		<float name="farClip" value="100.0"/>
		<float name="apertureRadius" value="0.02"/>
		<float name="focusDistance" value="3.4"/>
		<integer name="width" value="48"/>
		<integer name="height" value="32"/>
		<transform name="toWorld">
			<lookat origin="0, 0, -3.4" target="0, 0, 0" up="0, 1, 0"/>
		</transform>
		<rfilter type="mitchell"/>
	</camera>
	<mesh type="obj">
		<string name="filename" value="room.obj"/>
		<bsdf type="diffuse">
			<color name="albedo" value="0.7, 0.7 0.6"/>
		</bsdf>
	</mesh>
	<mesh type="obj">
		<string name="filename" value="sphere.obj"/>
		<transform name="toWorld">
			<scale value="0.45 0.45 0.45"/>
			<rotate angle="30" axis="0 1 0"/>
			<translate value="-0.3 -0.55 0.2"/>
		</transform>
		<bsdf type="normalmap">
			<texture type="imagetexture">
				<string name="filename" value="nmap.png"/>
				<string name="colorspace" value="linear"/>
				<float name="scale" value="3.0"/>
			</texture>
			<bsdf type="kazenstandard">
				<texture type="blend" id="baseColor">
					<string name="blendmode" value="mix"/>
					<texture type="constanttexture" id="mask"><color name="color" value="0.25 0.25 0.25"/></texture>
					<texture type="imagetexture" id="input1"><string name="filename" value="albedo.png"/></texture>
					<texture type="constanttexture" id="input2"><color name="color" value="0.871 0.376 0.0"/></texture>
				</texture>
				<texture type="colorramp" id="roughness">
					<float name="min" value="0.2"/>
					<float name="max" value="0.5"/>
					<texture type="imagetexture"><string name="filename" value="albedo.png"/><string name="colorspace" value="linear"/></texture>
				</texture>
				<texture type="constanttexture" id="metallic"><color name="color" value="0.1 0 0"/></texture>
				<float name="clearcoat" value="0.5"/>
				<float name="sheen" value="0.2"/>
			</bsdf>
		</bsdf>
	</mesh>
	<mesh type="obj">
		<string name="filename" value="light.obj"/>
		<light type="area">
			<color name="color" value="1.0 0.9 0.7"/>
			<float name="intensity" value="12.0"/>
		</light>
	</mesh>
	<texture type="background" id="background">
		<float name="intensity" value="0.5"/>
		<texture type="constanttexture"><color name="color" value="0.2 0.3 0.5"/></texture>
	</texture>
</scene>
"""


def make(dst=None):
    dst = dst or os.path.join(HERE, "box")
    os.makedirs(dst, exist_ok=True)
    open(os.path.join(dst, "room.obj"), "w").write(ROOM)
    open(os.path.join(dst, "light.obj"), "w").write(LIGHT)
    sphere_obj(os.path.join(dst, "sphere.obj"))
    rng = np.random.default_rng(3)
    write_png(os.path.join(dst, "albedo.png"), rng.integers(30, 230, (16, 32, 3), dtype=np.uint8))
    nm = np.zeros((8, 8, 3), np.uint8)
    nm[..., 0] = rng.integers(100, 156, (8, 8)); nm[..., 1] = rng.integers(100, 156, (8, 8)); nm[..., 2] = 245
    write_png(os.path.join(dst, "nmap.png"), nm)
    open(os.path.join(dst, "box.xml"), "w").write(XML)
    return os.path.join(dst, "box.xml")


if __name__ == "__main__":
    print(make())
