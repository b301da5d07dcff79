// instancefilemap.hpp -- the reference's ".inst" reader (/root/reference/utilities/instancefilemap.hpp:13-82)
// without Boost: lines "Key ? value", both sides trimmed; a line without '?' maps to itself.
#pragma once

#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

class InstanceFileMap {
public:
	InstanceFileMap() {}
	explicit InstanceFileMap(const std::string &instance) {
		std::ifstream file(instance.c_str());
		if (!file.is_open()) {
			fprintf(stderr, "can't open instance file: %s\n", instance.c_str());
			exit(1);
		}
		std::string line;
		while (std::getline(file, line)) {
			const std::size_t d = line.find('?');
			if (d == std::string::npos) map[line] = line;
			else map[trim(line.substr(0, d))] = trim(line.substr(d + 1));
		}
	}
	void set(const std::string &key, const std::string &value) { map[key] = value; }
	bool exists(const std::string &key) const { return map.find(key) != map.end(); }
	const std::string &value(const std::string &key) const {
		if (!exists(key)) {
			fprintf(stderr, "Key \"%s\" not bound\n", key.c_str());
			exit(1);
		}
		return map.at(key);
	}
	std::vector<std::string> valueList(const std::string &key, const std::string delim = " ") const {
		std::vector<std::string> list;
		const std::string &s = value(key);
		std::size_t b = 0;
		while (b <= s.size()) {
			const std::size_t e = s.find_first_of(delim, b);
			const std::string tok = s.substr(b, e == std::string::npos ? std::string::npos : e - b);
			if (!tok.empty()) list.push_back(tok);
			if (e == std::string::npos) break;
			b = e + 1;
		}
		return list;
	}

private:
	static std::string trim(const std::string &s) {
		const char *ws = " \t\r\n";
		const std::size_t b = s.find_first_not_of(ws);
		if (b == std::string::npos) return "";
		return s.substr(b, s.find_last_not_of(ws) - b + 1);
	}
	std::unordered_map<std::string, std::string> map;
};
