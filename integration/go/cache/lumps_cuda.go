// lumps_cuda.go -- cache.LumpCache (cache/bsp.go:21-47) as the C view the library's BSP-side entry points take
// (vrad_bsp_lumps, include/vrad_bsp.h).  The galaco/bsp primitive structs are Go structs, not guaranteed to share the
// on-disk layout, so every lump is marshalled once into C memory (C.malloc, freed by Free): the C side never holds
// a Go pointer (cgo rule).  SOURCE ONLY (no Go toolchain in the build image).
//
//go:build cuda

package cache

/*
#cgo CFLAGS:  -I${SRCDIR}/../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../vrad_b200/_lib -lvradcuda
#include <stdlib.h>
#include "vrad_bsp.h"
*/
import "C"

import "unsafe"

// CLumps owns the marshalled copies.
type CLumps struct {
	L    C.vrad_bsp_lumps
	bufs []unsafe.Pointer
}

func (c *CLumps) alloc(n int, size uintptr) unsafe.Pointer {
	if n == 0 {
		return nil
	}
	p := C.calloc(C.size_t(n), C.size_t(size))
	c.bufs = append(c.bufs, p)
	return p
}

func (c *CLumps) Free() {
	for _, p := range c.bufs {
		C.free(p)
	}
	c.bufs = nil
}

// BuildCLumps marshals the lumps the path consumes.  targetFaces = *cache.GetTargetFaces() (LDR or HDR faces,
// cmd/tasks/loadbsp/main.go:80-89).
func BuildCLumps() *CLumps {
	lc := GetLumpCache()
	faces := *GetTargetFaces()
	c := &CLumps{}
	L := &c.L

	L.n_planes = C.int32_t(len(lc.Planes))
	planes := (*[1 << 26]C.vrad_dplane)(c.alloc(len(lc.Planes), C.sizeof_vrad_dplane))
	for i, p := range lc.Planes {
		planes[i].normal[0], planes[i].normal[1], planes[i].normal[2] = C.float(p.Normal[0]), C.float(p.Normal[1]), C.float(p.Normal[2])
		planes[i].dist, planes[i]._type = C.float(p.Distance), C.int32_t(p.AxisType)
	}
	L.planes = &planes[0]

	L.n_vertexes = C.int32_t(len(lc.Vertexes))
	verts := (*[1 << 28]C.float)(c.alloc(3*len(lc.Vertexes), 4))
	for i, v := range lc.Vertexes {
		verts[3*i], verts[3*i+1], verts[3*i+2] = C.float(v[0]), C.float(v[1]), C.float(v[2])
	}
	L.vertexes3 = &verts[0]

	L.n_edges = C.int32_t(len(lc.Edges))
	edges := (*[1 << 26]C.vrad_dedge)(c.alloc(len(lc.Edges), C.sizeof_vrad_dedge))
	for i, e := range lc.Edges {
		edges[i].v[0], edges[i].v[1] = C.uint16_t(e[0]), C.uint16_t(e[1])
	}
	L.edges = &edges[0]

	L.n_surfedges = C.int32_t(len(lc.SurfEdges))
	se := (*[1 << 28]C.int32_t)(c.alloc(len(lc.SurfEdges), 4))
	for i, v := range lc.SurfEdges {
		se[i] = C.int32_t(v)
	}
	L.surfedges = &se[0]

	L.n_faces = C.int32_t(len(faces))
	cf := (*[1 << 24]C.vrad_dface)(c.alloc(len(faces), C.sizeof_vrad_dface))
	for i := range faces {
		f := &faces[i]
		cf[i].planenum, cf[i].side, cf[i].on_node = C.uint16_t(f.Planenum), C.uint8_t(f.Side), C.uint8_t(f.OnNode)
		cf[i].firstedge, cf[i].numedges = C.int32_t(f.FirstEdge), C.int16_t(f.NumEdges)
		cf[i].texinfo, cf[i].dispinfo, cf[i].fog_volume = C.int16_t(f.TexInfo), C.int16_t(f.DispInfo), C.int16_t(f.SurfaceFogVolumeID)
		for k := 0; k < 4; k++ {
			cf[i].styles[k] = C.uint8_t(f.Styles[k])
		}
		cf[i].lightofs, cf[i].area = C.int32_t(f.Lightofs), C.float(f.Area)
		for k := 0; k < 2; k++ {
			cf[i].lm_mins[k], cf[i].lm_size[k] = C.int32_t(f.LightmapTextureMinsInLuxels[k]), C.int32_t(f.LightmapTextureSizeInLuxels[k])
		}
		cf[i].orig_face, cf[i].smoothing_groups = C.int32_t(f.OrigFace), C.uint32_t(f.SmoothingGroups)
	}
	L.faces = &cf[0]

	L.n_texinfo = C.int32_t(len(lc.TexInfo))
	ti := (*[1 << 22]C.vrad_texinfo)(c.alloc(len(lc.TexInfo), C.sizeof_vrad_texinfo))
	for i := range lc.TexInfo {
		t := &lc.TexInfo[i]
		for a := 0; a < 2; a++ {
			for b := 0; b < 4; b++ {
				ti[i].texture_vecs[a][b] = C.float(t.TextureVecsTexelsPerWorldUnits[a][b])
				ti[i].lightmap_vecs[a][b] = C.float(t.LightmapVecsLuxelsPerWorldUnits[a][b])
			}
		}
		ti[i].flags, ti[i].texdata = C.int32_t(t.Flags), C.int32_t(t.TexData)
	}
	L.texinfo = &ti[0]

	L.n_texdata = C.int32_t(len(lc.TexData))
	td := (*[1 << 22]C.vrad_dtexdata)(c.alloc(len(lc.TexData), C.sizeof_vrad_dtexdata))
	for i := range lc.TexData {
		t := &lc.TexData[i]
		td[i].reflectivity[0], td[i].reflectivity[1], td[i].reflectivity[2] = C.float(t.Reflectivity[0]), C.float(t.Reflectivity[1]), C.float(t.Reflectivity[2])
		td[i].name_id, td[i].width, td[i].height = C.int32_t(t.NameStringTableID), C.int32_t(t.Width), C.int32_t(t.Height)
	}
	L.texdata = &td[0]

	L.n_models = C.int32_t(len(lc.Models))
	md := (*[1 << 20]C.vrad_dmodel)(c.alloc(len(lc.Models), C.sizeof_vrad_dmodel))
	for i := range lc.Models {
		m := &lc.Models[i]
		for k := 0; k < 3; k++ {
			md[i].mins[k], md[i].maxs[k], md[i].origin[k] = C.float(m.Mins[k]), C.float(m.Maxs[k]), C.float(m.Origin[k])
		}
		md[i].headnode, md[i].firstface, md[i].numfaces = C.int32_t(m.HeadNode), C.int32_t(m.FirstFace), C.int32_t(m.NumFaces)
	}
	L.models = &md[0]

	L.n_nodes = C.int32_t(len(lc.Nodes))
	nd := (*[1 << 24]C.vrad_dnode)(c.alloc(len(lc.Nodes), C.sizeof_vrad_dnode))
	for i := range lc.Nodes {
		n := &lc.Nodes[i]
		nd[i].planenum = C.int32_t(n.PlaneNum)
		nd[i].children[0], nd[i].children[1] = C.int32_t(n.Children[0]), C.int32_t(n.Children[1])
		for k := 0; k < 3; k++ {
			nd[i].mins[k], nd[i].maxs[k] = C.int16_t(n.Mins[k]), C.int16_t(n.Maxs[k])
		}
		nd[i].firstface, nd[i].numfaces, nd[i].area = C.uint16_t(n.FirstFace), C.uint16_t(n.NumFaces), C.int16_t(n.Area)
	}
	L.nodes = &nd[0]

	L.n_leafs = C.int32_t(len(lc.Leafs))
	lf := (*[1 << 24]C.vrad_dleaf)(c.alloc(len(lc.Leafs), C.sizeof_vrad_dleaf))
	for i := range lc.Leafs {
		l := &lc.Leafs[i]
		lf[i].contents, lf[i].cluster = C.int32_t(l.Contents), C.int16_t(l.Cluster)
		lf[i].area_flags = C.int16_t(int(l.Area())&0x1ff | int(l.Flags())<<9) // the 9:7 bit field of dleaf_t
		for k := 0; k < 3; k++ {
			lf[i].mins[k], lf[i].maxs[k] = C.int16_t(l.Mins[k]), C.int16_t(l.Maxs[k])
		}
		lf[i].firstleafface, lf[i].numleaffaces = C.uint16_t(l.FirstLeafFace), C.uint16_t(l.NumLeafFaces)
		lf[i].firstleafbrush, lf[i].numleafbrushes = C.uint16_t(l.FirstLeafBrush), C.uint16_t(l.NumLeafBrushes)
		lf[i].leaf_water_data = C.int16_t(l.LeafWaterDataID)
	}
	L.leafs = &lf[0]

	u16 := func(src []uint16) *C.uint16_t {
		dst := (*[1 << 28]C.uint16_t)(c.alloc(len(src), 2))
		for i, v := range src {
			dst[i] = C.uint16_t(v)
		}
		if len(src) == 0 {
			return nil
		}
		return &dst[0]
	}
	L.n_leaffaces, L.leaffaces = C.int32_t(len(lc.LeafFaces)), u16(lc.LeafFaces)
	L.n_leafbrushes, L.leafbrushes = C.int32_t(len(lc.LeafBrushes)), u16(lc.LeafBrushes)

	L.n_brushes = C.int32_t(len(lc.Brushes))
	br := (*[1 << 24]C.vrad_dbrush)(c.alloc(len(lc.Brushes), C.sizeof_vrad_dbrush))
	for i := range lc.Brushes {
		b := &lc.Brushes[i]
		br[i].firstside, br[i].numsides, br[i].contents = C.int32_t(b.FirstSide), C.int32_t(b.NumSides), C.int32_t(b.Contents)
	}
	L.brushes = &br[0]

	L.n_brushsides = C.int32_t(len(lc.BrushSides))
	bs := (*[1 << 24]C.vrad_dbrushside)(c.alloc(len(lc.BrushSides), C.sizeof_vrad_dbrushside))
	for i := range lc.BrushSides {
		s := &lc.BrushSides[i]
		bs[i].planenum, bs[i].texinfo, bs[i].dispinfo, bs[i].bevel = C.uint16_t(s.PlaneNum), C.int16_t(s.TexInfo), C.int16_t(s.DispInfo), C.int16_t(s.Bevel)
	}
	L.brushsides = &bs[0]

	L.n_areas = C.int32_t(len(lc.Areas))
	L.vis_len = C.int64_t(len(lc.VisDataRaw))
	if len(lc.VisDataRaw) > 0 {
		L.visdata = (*C.uint8_t)(C.CBytes(lc.VisDataRaw))
		c.bufs = append(c.bufs, unsafe.Pointer(L.visdata))
	}
	if rc := C.vrad_bsp_validate(L); rc != 0 { // every cross-lump index, once; the library's input functions assume it passed
		panic("vrad_bsp_validate: " + C.GoString(C.vrad_last_error()))
	}
	return c
}
