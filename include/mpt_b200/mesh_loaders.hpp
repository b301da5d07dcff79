// mesh_loaders.hpp -- SURVEY 8f N3: mesh ingestion without Assimp.
//
// The reference loads meshes with Assimp, `importer.ReadFile(pFile, 0)` -- no post-processing flags
// (/root/reference/utilities/assimp_mesh_loader.hpp:11-18) -- and hands FCL one (vertices, triangles) pair per
// aiMesh (`get`, :20-60).  These readers reproduce what that yields for the formats of the bundled fixtures:
//   COLLADA (.dae): one sub-mesh per <triangles> group (and per <polylist> / <polygons> group, of which only the
//                   three-vertex faces count) of every geometry that the visual scene instantiates,
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

// three fresh vertices per face, faces (0,1,2), (3,4,5), ...  Returns false if a face index points past the
// vertex array (malformed file): mesh files are untrusted input.
inline bool unindex(const std::vector<float> &pos, const std::vector<uint32_t> &faces, SubMesh &m) {
	const size_t nv = pos.size() / 3;
	for (size_t i = 0; i < faces.size(); ++i) {
		const size_t v = faces[i];
		if (v >= nv) return false;
		for (int k = 0; k < 3; ++k) m.xyz.push_back((double)pos[3 * v + k]);
		m.tri.push_back((uint32_t)i);
	}
	return true;
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
			if (stride < 1 || stride > 64) return false;  // (a stride of 0 would never advance)
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
			for (size_t i = 0; i + (size_t)stride <= vals.size(); i += (size_t)stride)
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
			// <triangles>; <polylist> / <polygons> faces with exactly three vertices (Assimp without the triangulate flag
			// keeps polygons as they are and AssimpMeshLoader::get only takes 3-index faces,
			// /root/reference/utilities/assimp_mesh_loader.hpp:50-53).  <lines> etc. have no such faces.
			const bool is_tri = prim->name == "triangles", is_plist = prim->name == "polylist", is_pgons = prim->name == "polygons";
			if (!is_tri && !is_plist && !is_pgons) continue;
			int nin = 1, voff = -1;
			const std::vector<float> *pos = nullptr;
			for (const auto &in : prim->kids) {
				if (in->name != "input") continue;
				const int off = atoi(in->get("offset").c_str());
				if (off < 0 || off > 255) return false;
				if (off + 1 > nin) nin = off + 1;
				if (in->get("semantic") == "VERTEX") {
					voff = off;
					auto it = verts.find(strip_hash(in->get("source")));
					if (it != verts.end()) pos = it->second;
				}
			}
			if (!pos || voff < 0) continue;
			auto read_indices = [nin, voff](const std::string &text, std::vector<uint32_t> &idx) {
				const char *c = text.c_str();
				char *end = nullptr;
				long col = 0;
				for (;;) {
					const long v = strtol(c, &end, 10);
					if (end == c) break;
					if (col % nin == voff) idx.push_back(v < 0 ? 0xffffffffu : (uint32_t)v);  // (negative: rejected by unindex)
					++col;
					c = end;
				}
			};
			std::vector<uint32_t> idx;
			if (is_pgons) {
				for (const auto &pp : prim->kids) {
					if (pp->name != "p") continue;
					std::vector<uint32_t> one;
					read_indices(pp->text, one);
					if (one.size() == 3) idx.insert(idx.end(), one.begin(), one.end());
				}
			} else {
				const Xml *p = child(*prim, "p");
				if (!p) continue;
				std::vector<uint32_t> all;
				read_indices(p->text, all);
				if (is_tri) {
					all.resize(all.size() / 3 * 3);
					idx.swap(all);
				} else {
					const Xml *vc = child(*prim, "vcount");
					if (!vc) continue;
					const char *c = vc->text.c_str();
					char *end = nullptr;
					size_t at = 0;
					for (;;) {
						const long n = strtol(c, &end, 10);
						if (end == c) break;
						c = end;
						if (n < 0 || at + (size_t)n > all.size()) return false;
						if (n == 3) idx.insert(idx.end(), all.begin() + (long)at, all.begin() + (long)at + 3);
						at += (size_t)n;
					}
				}
			}
			if (idx.empty()) continue;
			SubMesh sm;
			if (!unindex(*pos, idx, sm)) return false;
			out.push_back(sm);
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
		bool bad;  // a chunk whose payload runs past its own end: the file is rejected
		uint16_t u16(size_t o) const { uint16_t v; memcpy(&v, d.data() + o, 2); return v; }
		uint32_t u32(size_t o) const { uint32_t v; memcpy(&v, d.data() + o, 4); return v; }
		std::string cstr(size_t o, size_t &next) const {
			size_t e = o;
			while (e < d.size() && d[e]) ++e;
			next = e + 1;
			return d.substr(o, e - o);
		}
		void walk(size_t off, size_t end, Obj *obj, std::string *matname) {
			while (off + 6 <= end && !bad) {
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
					if (body + 2 > cend) { bad = true; return; }
					const uint16_t n = u16(body);
					if (body + 2 + sizeof(float) * 3 * (size_t)n > cend) { bad = true; return; }  // payload must fit the chunk
					obj->verts.resize(3 * (size_t)n);
					memcpy(obj->verts.data(), d.data() + body + 2, sizeof(float) * 3 * n);
				} else if (id == 0x4120 && obj) {
					if (body + 2 > cend) { bad = true; return; }
					const uint16_t n = u16(body);
					if (body + 2 + 8 * (size_t)n > cend) { bad = true; return; }
					for (uint16_t f = 0; f < n; ++f)
						for (int k = 0; k < 3; ++k) obj->faces.push_back(u16(body + 2 + 8 * (size_t)f + 2 * k));
					walk(body + 2 + 8 * (size_t)n, cend, obj, matname);
				} else if (id == 0x4130 && obj) {
					size_t nx;
					const std::string m = cstr(body, nx);
					if (nx + 2 > cend) { bad = true; return; }
					const uint16_t n = u16(nx);
					if (nx + 2 + 2 * (size_t)n > cend) { bad = true; return; }
					std::vector<uint32_t> ids;
					for (uint16_t f = 0; f < n; ++f) ids.push_back(u16(nx + 2 + 2 * (size_t)f));
					obj->facemat.push_back(std::make_pair(m, ids));
				}
				off = cend;
			}
		}
	} parser{d, materials, objects, false};
	parser.walk(0, d.size(), nullptr, nullptr);
	if (parser.bad) return false;
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
			if (!idx.empty()) {
				SubMesh sm;
				if (!unindex(o.verts, idx, sm)) return false;
				out.push_back(sm);
			}
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
	// the header counts are checked against what is left of the file before anything is allocated
	f.seekg(0, std::ios::end);
	const unsigned long long file_size = (unsigned long long)f.tellg();
	f.seekg(12, std::ios::beg);
	unsigned long long left = file_size >= 12 ? file_size - 12 : 0;
	if ((unsigned long long)nsub * 8ull > left) return false;
	out.resize(nsub);
	for (auto &m : out) {
		uint32_t nv = 0, nt = 0;
		f.read(reinterpret_cast<char *>(&nv), 4);
		f.read(reinterpret_cast<char *>(&nt), 4);
		if (!f || left < 8) return false;
		left -= 8;
		const unsigned long long need = 24ull * nv + 12ull * nt;
		if (need > left) return false;
		left -= need;
		m.xyz.resize(3 * (size_t)nv);
		m.tri.resize(3 * (size_t)nt);
		f.read(reinterpret_cast<char *>(m.xyz.data()), sizeof(double) * m.xyz.size());
		f.read(reinterpret_cast<char *>(m.tri.data()), sizeof(uint32_t) * m.tri.size());
		if (!f) return false;
		for (uint32_t i : m.tri)
			if (i >= nv) return false;
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
