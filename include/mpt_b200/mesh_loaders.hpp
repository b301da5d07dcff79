// mesh_loaders.hpp -- SURVEY 8f N3: mesh ingestion without Assimp.
//
// The reference loads meshes with Assimp, `importer.ReadFile(pFile, 0)` -- no post-processing flags
// (/root/reference/utilities/assimp_mesh_loader.hpp:11-18) -- and hands FCL one (vertices, triangles) pair per
// aiMesh (`get`, :20-60).  These readers reproduce what that yields for the formats of the bundled fixtures:
//   COLLADA (.dae): one sub-mesh per <triangles> group of every geometry that the visual scene instantiates,
//                   in Assimp's ColladaLoader::BuildHierarchy order (real child nodes, then <instance_node>s,
//                   then the node's own <instance_geometry>), each geometry once; vertices un-indexed (three
//                   fresh vertices per face), raw mesh-local float32 coordinates (node matrices, <unit>,
//                   up_axis ignored); <lines> groups have no 3-index faces, their sub-meshes end up without
//                   triangles and are skipped by the mesh handlers, so they are not emitted.
//   3D Studio (.3ds): one sub-mesh per (object, material) in material order, default material last.
//   .mptm: the flat format written by tools/convert_meshes.py and by write_mptm() below.
// tools/convert_meshes.py is the same logic in Python; tests/test_host_library.py checks both agree byte for
// byte on the reference's fixtures.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

namespace mpt_b200 {

struct SubMesh {
	std::vector<double> xyz;     // [nv*3]
	std::vector<uint32_t> tri;   // [nt*3]
};

namespace detail {

inline bool read_file(const std::string &path, std::string &out) {
	std::ifstream f(path.c_str(), std::ios::binary);
	if (!f.is_open()) return false;
	std::ostringstream ss;
	ss << f.rdbuf();
	out = ss.str();
	return true;
}

// three fresh vertices per face, faces (0,1,2), (3,4,5), ...
inline SubMesh unindex(const std::vector<float> &pos, const std::vector<uint32_t> &faces) {
	SubMesh m;
	for (size_t i = 0; i < faces.size(); ++i) {
		const size_t v = faces[i];
		for (int k = 0; k < 3; ++k) m.xyz.push_back((double)pos[3 * v + k]);
		m.tri.push_back((uint32_t)i);
	}
	return m;
}

// ---- a small XML reader: elements, attributes, text; comments / processing instructions skipped ----------
struct Xml {
	std::string name, text;
	std::map<std::string, std::string> attr;
	std::vector<std::unique_ptr<Xml> > kids;
	const std::string &get(const std::string &k) const {
		static const std::string empty;
		auto it = attr.find(k);
		return it == attr.end() ? empty : it->second;
	}
};

inline bool parse_xml(const std::string &s, Xml &root) {
	std::vector<Xml *> stack(1, &root);
	size_t i = 0;
	while (i < s.size()) {
		if (s[i] != '<') {
			const size_t j = s.find('<', i);
			stack.back()->text.append(s, i, (j == std::string::npos ? s.size() : j) - i);
			i = j == std::string::npos ? s.size() : j;
			continue;
		}
		if (s.compare(i, 4, "<!--") == 0) {
			const size_t j = s.find("-->", i);
			if (j == std::string::npos) return false;
			i = j + 3;
			continue;
		}
		if (s.compare(i, 2, "<?") == 0 || s.compare(i, 2, "<!") == 0) {
			const size_t j = s.find('>', i);
			if (j == std::string::npos) return false;
			i = j + 1;
			continue;
		}
		const size_t close = s.find('>', i);
		if (close == std::string::npos) return false;
		if (s[i + 1] == '/') {
			if (stack.size() > 1) stack.pop_back();
			i = close + 1;
			continue;
		}
		const bool selfclose = s[close - 1] == '/';
		const std::string tag = s.substr(i + 1, close - i - 1 - (selfclose ? 1 : 0));
		std::unique_ptr<Xml> e(new Xml);
		size_t p = 0;
		while (p < tag.size() && !isspace((unsigned char)tag[p])) ++p;
		e->name = tag.substr(0, p);
		const size_t colon = e->name.find(':');
		if (colon != std::string::npos) e->name = e->name.substr(colon + 1);
		while (p < tag.size()) {
			while (p < tag.size() && isspace((unsigned char)tag[p])) ++p;
			const size_t eq = tag.find('=', p);
			if (eq == std::string::npos) break;
			std::string key = tag.substr(p, eq - p);
			while (!key.empty() && isspace((unsigned char)key.back())) key.pop_back();
			size_t q = eq + 1;
			while (q < tag.size() && isspace((unsigned char)tag[q])) ++q;
			if (q >= tag.size()) break;
			const char quote = tag[q];
			const size_t endq = tag.find(quote, q + 1);
			if (endq == std::string::npos) break;
			e->attr[key] = tag.substr(q + 1, endq - q - 1);
			p = endq + 1;
		}
		Xml *raw = e.get();
		stack.back()->kids.push_back(std::move(e));
		if (!selfclose) stack.push_back(raw);
		i = close + 1;
	}
	return true;
}

inline void find_all(const Xml &n, const std::string &name, std::vector<const Xml *> &out) {
	for (const auto &k : n.kids) {
		if (k->name == name) out.push_back(k.get());
		find_all(*k, name, out);
	}
}

inline const Xml *child(const Xml &n, const std::string &name) {
	for (const auto &k : n.kids)
		if (k->name == name) return k.get();
	return nullptr;
}

inline std::string strip_hash(const std::string &u) { return !u.empty() && u[0] == '#' ? u.substr(1) : u; }

inline bool load_dae(const std::string &path, std::vector<SubMesh> &out) {
	std::string buf;
	if (!read_file(path, buf)) return false;
	Xml doc;
	if (!parse_xml(buf, doc)) return false;
	std::map<std::string, const Xml *> geoms, nodes;
	std::vector<const Xml *> tmp;
	find_all(doc, "geometry", tmp);
	for (const Xml *g : tmp) geoms[g->get("id")] = g;
	tmp.clear();
	find_all(doc, "node", tmp);
	for (const Xml *n : tmp)
		if (!n->get("id").empty()) nodes[n->get("id")] = n;
	// instantiation order
	std::vector<std::string> order;
	struct Walk {
		const std::map<std::string, const Xml *> &nodes;
		std::vector<std::string> &order;
		void visit(const Xml &n) {
			for (const auto &k : n.kids)
				if (k->name == "node") visit(*k);
			for (const auto &k : n.kids)
				if (k->name == "instance_node") {
					auto it = nodes.find(strip_hash(k->get("url")));
					if (it != nodes.end()) visit(*it->second);
				}
			for (const auto &k : n.kids)
				if (k->name == "instance_geometry") {
					const std::string id = strip_hash(k->get("url"));
					bool seen = false;
					for (const auto &o : order) seen = seen || o == id;
					if (!seen) order.push_back(id);
				}
		}
	} walk{nodes, order};
	tmp.clear();
	find_all(doc, "visual_scene", tmp);
	for (const Xml *vs : tmp)
		for (const auto &k : vs->kids)
			if (k->name == "node") walk.visit(*k);

	for (const std::string &gid : order) {
		auto git = geoms.find(gid);
		if (git == geoms.end()) continue;
		const Xml *mesh = child(*git->second, "mesh");
		if (!mesh) continue;
		std::map<std::string, std::vector<float> > sources;   // id -> xyz triples
		for (const auto &k : mesh->kids) {
			if (k->name != "source") continue;
			const Xml *fa = child(*k, "float_array");
			if (!fa) continue;
			int stride = 3;
			const Xml *tc = child(*k, "technique_common");
			const Xml *acc = tc ? child(*tc, "accessor") : nullptr;
			if (acc && !acc->get("stride").empty()) stride = atoi(acc->get("stride").c_str());
			std::vector<float> vals;
			const char *c = fa->text.c_str();
			char *end = nullptr;
			for (;;) {
				const double d = strtod(c, &end);
				if (end == c) break;
				vals.push_back((float)d);
				c = end;
			}
			std::vector<float> xyz;
			for (size_t i = 0; i + stride <= vals.size(); i += stride)
				for (int a = 0; a < 3; ++a) xyz.push_back(a < stride ? vals[i + a] : 0.f);
			sources[k->get("id")] = xyz;
		}
		std::map<std::string, const std::vector<float> *> verts;
		for (const auto &k : mesh->kids) {
			if (k->name != "vertices") continue;
			for (const auto &in : k->kids)
				if (in->name == "input" && in->get("semantic") == "POSITION") {
					auto it = sources.find(strip_hash(in->get("source")));
					if (it != sources.end()) verts[k->get("id")] = &it->second;
				}
		}
		for (const auto &prim : mesh->kids) {
			if (prim->name != "triangles") continue;   // <lines> etc.: no triangle faces -> skipped sub-mesh
			int nin = 1, voff = -1;
			const std::vector<float> *pos = nullptr;
			for (const auto &in : prim->kids) {
				if (in->name != "input") continue;
				const int off = atoi(in->get("offset").c_str());
				if (off + 1 > nin) nin = off + 1;
				if (in->get("semantic") == "VERTEX") {
					voff = off;
					auto it = verts.find(strip_hash(in->get("source")));
					if (it != verts.end()) pos = it->second;
				}
			}
			const Xml *p = child(*prim, "p");
			if (!p || !pos || voff < 0) continue;
			std::vector<uint32_t> idx;
			const char *c = p->text.c_str();
			char *end = nullptr;
			long col = 0;
			for (;;) {
				const long v = strtol(c, &end, 10);
				if (end == c) break;
				if (col % nin == voff) idx.push_back((uint32_t)v);
				++col;
				c = end;
			}
			idx.resize(idx.size() / 3 * 3);
			out.push_back(unindex(*pos, idx));
		}
	}
	return true;
}

inline bool load_3ds(const std::string &path, std::vector<SubMesh> &out) {
	std::string d;
	if (!read_file(path, d)) return false;
	struct Obj {
		std::vector<float> verts;
		std::vector<uint32_t> faces;
		std::vector<std::pair<std::string, std::vector<uint32_t> > > facemat;
	};
	std::vector<std::string> materials;
	std::vector<Obj> objects;
	struct Parser {
		const std::string &d;
		std::vector<std::string> &materials;
		std::vector<Obj> &objects;
		uint16_t u16(size_t o) const { uint16_t v; memcpy(&v, d.data() + o, 2); return v; }
		uint32_t u32(size_t o) const { uint32_t v; memcpy(&v, d.data() + o, 4); return v; }
		std::string cstr(size_t o, size_t &next) const {
			size_t e = o;
			while (e < d.size() && d[e]) ++e;
			next = e + 1;
			return d.substr(o, e - o);
		}
		void walk(size_t off, size_t end, Obj *obj, std::string *matname) {
			while (off + 6 <= end) {
				const uint16_t id = u16(off);
				const uint32_t len = u32(off + 2);
				const size_t body = off + 6, cend = off + len;
				if (len < 6 || cend > d.size()) return;
				if (id == 0x4D4D || id == 0x3D3D) {
					walk(body, cend, obj, matname);
				} else if (id == 0xAFFF) {
					std::string name;
					walk(body, cend, nullptr, &name);
					materials.push_back(name);
				} else if (id == 0xA000 && matname) {
					size_t nx;
					*matname = cstr(body, nx);
				} else if (id == 0x4000) {
					size_t nx;
					cstr(body, nx);
					objects.emplace_back();
					walk(nx, cend, &objects.back(), nullptr);
				} else if (id == 0x4100) {
					walk(body, cend, obj, matname);
				} else if (id == 0x4110 && obj) {
					const uint16_t n = u16(body);
					obj->verts.resize(3 * (size_t)n);
					memcpy(obj->verts.data(), d.data() + body + 2, sizeof(float) * 3 * n);
				} else if (id == 0x4120 && obj) {
					const uint16_t n = u16(body);
					for (uint16_t f = 0; f < n; ++f)
						for (int k = 0; k < 3; ++k) obj->faces.push_back(u16(body + 2 + 8 * (size_t)f + 2 * k));
					walk(body + 2 + 8 * (size_t)n, cend, obj, matname);
				} else if (id == 0x4130 && obj) {
					size_t nx;
					const std::string m = cstr(body, nx);
					const uint16_t n = u16(nx);
					std::vector<uint32_t> ids;
					for (uint16_t f = 0; f < n; ++f) ids.push_back(u16(nx + 2 + 2 * (size_t)f));
					obj->facemat.push_back(std::make_pair(m, ids));
				}
				off = cend;
			}
		}
	} parser{d, materials, objects};
	parser.walk(0, d.size(), nullptr, nullptr);
	for (const Obj &o : objects) {
		const size_t nf = o.faces.size() / 3;
		if (nf == 0 || o.verts.empty()) continue;            // Assimp drops empty objects
		std::vector<int> fm(nf, -1);
		for (const auto &pm : o.facemat) {
			int mi = -1;
			for (size_t k = 0; k < materials.size(); ++k)
				if (materials[k] == pm.first) { mi = (int)k; break; }
			for (uint32_t f : pm.second)
				if (f < nf) fm[f] = mi;
		}
		for (int mi = 0; mi <= (int)materials.size(); ++mi) {   // material order, default material (-1) last
			const int want = mi < (int)materials.size() ? mi : -1;
			std::vector<uint32_t> idx;
			for (size_t f = 0; f < nf; ++f)
				if (fm[f] == want)
					for (int k = 0; k < 3; ++k) idx.push_back(o.faces[3 * f + k]);
			if (!idx.empty()) out.push_back(unindex(o.verts, idx));
		}
	}
	return true;
}

inline bool load_mptm(const std::string &path, std::vector<SubMesh> &out) {
	std::ifstream f(path.c_str(), std::ios::binary);
	char magic[4];
	uint32_t ver = 0, nsub = 0;
	if (!f.is_open() || !f.read(magic, 4) || std::memcmp(magic, "MPTM", 4) != 0 ||
	    !f.read(reinterpret_cast<char *>(&ver), 4) || ver != 1 || !f.read(reinterpret_cast<char *>(&nsub), 4))
		return false;
	out.resize(nsub);
	for (auto &m : out) {
		uint32_t nv = 0, nt = 0;
		f.read(reinterpret_cast<char *>(&nv), 4);
		f.read(reinterpret_cast<char *>(&nt), 4);
		m.xyz.resize(3 * (size_t)nv);
		m.tri.resize(3 * (size_t)nt);
		f.read(reinterpret_cast<char *>(m.xyz.data()), sizeof(double) * m.xyz.size());
		f.read(reinterpret_cast<char *>(m.tri.data()), sizeof(uint32_t) * m.tri.size());
		if (!f) return false;
	}
	return true;
}

inline bool ends_with(const std::string &s, const char *suffix) {
	const size_t n = strlen(suffix);
	if (s.size() < n) return false;
	for (size_t i = 0; i < n; ++i)
		if (tolower((unsigned char)s[s.size() - n + i]) != suffix[i]) return false;
	return true;
}

}  // namespace detail

inline bool write_mptm(const std::string &path, const std::vector<SubMesh> &meshes) {
	std::ofstream f(path.c_str(), std::ios::binary);
	if (!f.is_open()) return false;
	const uint32_t ver = 1, n = (uint32_t)meshes.size();
	f.write("MPTM", 4);
	f.write(reinterpret_cast<const char *>(&ver), 4);
	f.write(reinterpret_cast<const char *>(&n), 4);
	for (const auto &m : meshes) {
		const uint32_t nv = (uint32_t)(m.xyz.size() / 3), nt = (uint32_t)(m.tri.size() / 3);
		f.write(reinterpret_cast<const char *>(&nv), 4);
		f.write(reinterpret_cast<const char *>(&nt), 4);
		f.write(reinterpret_cast<const char *>(m.xyz.data()), sizeof(double) * m.xyz.size());
		f.write(reinterpret_cast<const char *>(m.tri.data()), sizeof(uint32_t) * m.tri.size());
	}
	return (bool)f;
}

// Stand-in for AssimpMeshLoader: one SubMesh per aiMesh, in Assimp's order; .dae, .3ds and .mptm.
class MeshFileLoader {
public:
	explicit MeshFileLoader(const std::string &file) : error(false) {
		bool ok;
		if (detail::ends_with(file, ".dae")) ok = detail::load_dae(file, meshes);
		else if (detail::ends_with(file, ".3ds")) ok = detail::load_3ds(file, meshes);
		else ok = detail::load_mptm(file, meshes);
		if (!ok) {
			fprintf(stderr, "%s: cannot read mesh (supported: .dae <triangles>, .3ds, .mptm)\n", file.c_str());
			error = true;
		}
	}
	std::vector<SubMesh> meshes;
	bool error;
};

}  // namespace mpt_b200
