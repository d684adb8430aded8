// OBJ / MTL ingestion, host half: the text parse and the vertex de-duplication of AR::Mesh(path)
// (reference src/mesh.cpp:300-415 parseModelFile, :65-220 loadMaterial / parseMaterialData), restated so that the arrays are the
// reference loader's arrays: vertex order = first occurrence of a (position, uv, normal) value triple, fan triangulation, one
// material group per `usemtl`. Tangents / bitangents are generated on the device afterwards (axr_tangents.cuh).
//
// The reference reads every line through std::istringstream: `>> float` takes the longest prefix that looks like a decimal
// floating-point number (no "inf", "nan" or hex), a failed extraction stores 0 and makes the rest of the line fail too;
// face corners are split at '/', each part goes through std::stoi (which throws on a part without digits: reported here as an
// error instead of terminating the process); OBJ's negative (relative) indices are not supported by the reference: a corner
// whose position index is < 1 or beyond the positions read SO FAR is dropped, out-of-range uv / normal indices mean zeros.
#pragma once
#include <cctype>
#include <cerrno>
#include <climits>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

namespace axr_obj {

struct Group { std::string name; uint64_t first_face, face_count; };
struct Parsed {
	std::vector<float> v8;       // position3, uv2, normal3 per unique vertex
	std::vector<uint32_t> idx;   // 3 per face
	std::vector<Group> groups;
	std::string error;
};

// Cursor over one line [p, e) behaving like an std::istringstream on that line
struct LineStream {
	const char* p;
	const char* e;
	bool failed = false;
	static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f' || c == '\n'; }
	void skip_ws() { while (p < e && is_space(*p)) ++p; }
	// operator>>(std::string&): false at end of line
	bool word(const char*& b, const char*& end) {
		if (failed) return false;
		skip_ws();
		if (p >= e) { failed = true; return false; }
		b = p;
		while (p < e && !is_space(*p)) ++p;
		end = p;
		return true;
	}
	// operator>>(float&): the characters num_get accepts ([+-]digits[.digits][e[+-]digits]), converted by strtof; a failure leaves 0
	float number() {
		if (failed) return 0.0f;
		skip_ws();
		const char* q = p;
		if (q < e && (*q == '+' || *q == '-')) ++q;
		const char* digits = q;
		while (q < e && isdigit((unsigned char)*q)) ++q;
		bool any = q > digits;
		if (q < e && *q == '.') {
			++q;
			const char* frac = q;
			while (q < e && isdigit((unsigned char)*q)) ++q;
			any = any || q > frac;
		}
		if (!any) { failed = true; return 0.0f; }
		if (q < e && (*q == 'e' || *q == 'E')) {
			const char* x = q + 1;
			if (x < e && (*x == '+' || *x == '-')) ++x;
			if (x < e && isdigit((unsigned char)*x)) {
				while (x < e && isdigit((unsigned char)*x)) ++x;
				q = x;
			}
		}
		char buf[128];
		size_t n = (size_t)(q - p);
		if (n >= sizeof buf) n = sizeof buf - 1;  // a number of more than 127 characters: the tail cannot change a binary32 value
		memcpy(buf, p, n);
		buf[n] = 0;
		p = q;
		return strtof(buf, nullptr);
	}
};

// std::stoi on [b, e): optional whitespace, sign, decimal digits; trailing characters are ignored. false = it would have thrown.
inline bool stoi_like(const char* b, const char* e, int& out) {
	while (b < e && LineStream::is_space(*b)) ++b;
	bool neg = false;
	if (b < e && (*b == '+' || *b == '-')) { neg = *b == '-'; ++b; }
	if (b >= e || !isdigit((unsigned char)*b)) return false;
	long long v = 0;
	while (b < e && isdigit((unsigned char)*b)) {
		v = v * 10 + (*b - '0');
		if (v > (long long)INT_MAX + 1) return false;  // std::out_of_range
		++b;
	}
	if (neg) v = -v;
	if (v > INT_MAX || v < INT_MIN) return false;
	out = (int)v;
	return true;
}

struct Key8 {
	float f[8];
	bool operator==(const Key8& o) const {  // Vertex::operator== (reference include/mesh.hpp:15-17): float ==, so -0 == +0 and NaN != NaN
		for (int i = 0; i < 8; ++i) if (!(f[i] == o.f[i])) return false;
		return true;
	}
};
struct Key8Hash {
	size_t operator()(const Key8& k) const {
		uint64_t h = 0x9E3779B97F4A7C15ull;
		for (int i = 0; i < 8; ++i) {
			uint32_t b;
			float v = k.f[i] == 0.0f ? 0.0f : k.f[i];  // -0 and +0 are equal: they have to hash alike
			memcpy(&b, &v, 4);
			h = (h ^ b) * 0x100000001B3ull;
			h ^= h >> 29;
		}
		return (size_t)h;
	}
};

inline bool parse_obj(const char* text, size_t len, Parsed& out) {
	std::vector<float> positions, uvs, normals;  // 3, 2, 3 floats each
	std::unordered_map<Key8, uint32_t, Key8Hash> unique;
	std::vector<uint32_t> corner;
	const char* p = text;
	const char* end = text + len;
	uint64_t line_no = 0;
	while (p < end) {
		const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
		const char* le = nl ? nl : end;
		++line_no;
		LineStream ls{p, le};
		p = nl ? nl + 1 : end;
		if (ls.p == ls.e) continue;
		const char *tb, *te;
		if (!ls.word(tb, te)) continue;
		const size_t tl = (size_t)(te - tb);
		if (tl == 1 && tb[0] == 'v') {
			const float x = ls.number(), y = ls.number(), z = ls.number();
			positions.push_back(x); positions.push_back(y); positions.push_back(z);
		} else if (tl == 2 && tb[0] == 'v' && tb[1] == 't') {
			const float u = ls.number(), v = ls.number();
			uvs.push_back(u); uvs.push_back(v);
		} else if (tl == 2 && tb[0] == 'v' && tb[1] == 'n') {
			const float x = ls.number(), y = ls.number(), z = ls.number();
			normals.push_back(x); normals.push_back(y); normals.push_back(z);
		} else if (tl == 6 && memcmp(tb, "usemtl", 6) == 0) {
			const char *nb, *ne;
			std::string name;
			if (ls.word(nb, ne)) name.assign(nb, ne);
			else if (!out.groups.empty()) name = out.groups.back().name;  // a failed extraction leaves currentMaterialName as it was
			const uint64_t nfaces = out.idx.size() / 3;
			if (!out.groups.empty()) out.groups.back().face_count = nfaces - out.groups.back().first_face;
			out.groups.push_back({name, nfaces, 0});
		} else if (tl == 1 && tb[0] == 'f') {
			corner.clear();
			const char *cb, *ce;
			while (ls.word(cb, ce)) {
				int part[3] = {-1, -1, -1};
				const char* s = cb;
				for (int k = 0; k < 3 && s <= ce; ++k) {
					const char* slash = (const char*)memchr(s, '/', (size_t)(ce - s));
					const char* pe = slash ? slash : ce;
					if (pe > s) {
						int v;
						if (!stoi_like(s, pe, v)) {
							out.error = "face index without digits at line " + std::to_string(line_no) + " (std::stoi throws in the reference)";
							return false;
						}
						part[k] = v - 1;
					}
					if (!slash) break;
					s = slash + 1;
				}
				const int vi = part[0], ti = part[1], ni = part[2];
				if (vi < 0 || (size_t)vi >= positions.size() / 3) continue;
				Key8 key;
				key.f[0] = positions[3 * (size_t)vi]; key.f[1] = positions[3 * (size_t)vi + 1]; key.f[2] = positions[3 * (size_t)vi + 2];
				const bool has_uv = ti >= 0 && (size_t)ti < uvs.size() / 2, has_n = ni >= 0 && (size_t)ni < normals.size() / 3;
				key.f[3] = has_uv ? uvs[2 * (size_t)ti] : 0.0f; key.f[4] = has_uv ? uvs[2 * (size_t)ti + 1] : 0.0f;
				key.f[5] = has_n ? normals[3 * (size_t)ni] : 0.0f; key.f[6] = has_n ? normals[3 * (size_t)ni + 1] : 0.0f;
				key.f[7] = has_n ? normals[3 * (size_t)ni + 2] : 0.0f;
				auto it = unique.find(key);
				uint32_t id;
				if (it == unique.end()) {
					id = (uint32_t)(out.v8.size() / 8);
					unique.emplace(key, id);
					out.v8.insert(out.v8.end(), key.f, key.f + 8);
				} else {
					id = it->second;
				}
				corner.push_back(id);
			}
			for (size_t i = 1; i + 1 < corner.size(); ++i) {  // fan (:387-395)
				out.idx.push_back(corner[0]); out.idx.push_back(corner[i]); out.idx.push_back(corner[i + 1]);
			}
		}
	}
	if (!out.groups.empty()) out.groups.back().face_count = out.idx.size() / 3 - out.groups.back().first_face;
	return true;
}

// ---- MTL (reference src/mesh.cpp:65-220): material names, Ns and the texture paths of the five slots the shaders read
struct MtlEntry { std::string name; float specular_exponent = 0.0f; std::string map[5]; bool has_map[5] = {false, false, false, false, false}; };

inline void parse_mtl(const char* text, size_t len, std::vector<MtlEntry>& out) {
	const char* p = text;
	const char* end = text + len;
	while (p < end) {
		const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
		const char* le = nl ? nl : end;
		LineStream ls{p, le};
		p = nl ? nl + 1 : end;
		if (ls.p == ls.e) continue;
		const char *tb, *te;
		if (!ls.word(tb, te)) continue;
		const std::string type(tb, te);
		if (type == "newmtl") {
			MtlEntry m;
			const char *nb, *ne;
			if (ls.word(nb, ne)) m.name.assign(nb, ne);
			out.push_back(m);
			continue;
		}
		if (out.empty()) continue;
		MtlEntry& cur = out.back();
		int slot = -1;  // axr_set_material order: diffuse, bump, metallic, roughness, ao
		if (type == "Ns") cur.specular_exponent = ls.number();
		else if (type == "map_Kd") slot = 0;
		else if (type == "map_Bump" || type == "bump" || type == "norm") slot = 1;
		else if (type == "map_Ks" || type == "refl") slot = 2;
		else if (type == "map_Ns") slot = 3;
		else if (type == "map_A0") slot = 4;
		if (slot >= 0) {
			// rest of the line without leading / trailing blanks and tabs (processTexturePathLine)
			const char* b = ls.p;
			const char* e2 = ls.e;
			while (b < e2 && (*b == ' ' || *b == '\t')) ++b;
			while (e2 > b && (e2[-1] == ' ' || e2[-1] == '\t')) --e2;
			cur.map[slot].assign(b, e2);
			cur.has_map[slot] = true;
		}
	}
}

}  // namespace axr_obj
