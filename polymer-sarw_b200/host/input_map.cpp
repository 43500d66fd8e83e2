#include "input_map.h"
#include <cstdlib>
#include <fstream>
#include <sstream>

namespace sarwhost {

void trim(std::string& s)
{
	const char* ws = " \t\r\n\v\f";
	size_t b = s.find_first_not_of(ws);
	if (b == std::string::npos) { s.clear(); return; }
	s = s.substr(b, s.find_last_not_of(ws) - b + 1);
}

void environmentSubstitute(std::string& s)
{
	std::string out;
	for (size_t i = 0; i < s.size();) {
		if (s[i] != '$' || i + 1 >= s.size()) { out += s[i++]; continue; }
		size_t b = i + 1, e; const bool braced = s[b] == '{';
		if (braced) { ++b; e = s.find('}', b); if (e == std::string::npos) { out += s[i++]; continue; } }
		else { e = b; while (e < s.size() && (isalnum((unsigned char)s[e]) || s[e] == '_')) ++e; }
		const std::string name = s.substr(b, e - b);
		if (name.empty()) { out += s[i++]; continue; }
		if (const char* val = getenv(name.c_str())) out += val;
		i = braced ? e + 1 : e;
	}
	s = out;
}

bool InputMap::load(const std::string& filename, std::string& error)
{
	std::ifstream ifs(filename.c_str());
	if (!ifs.is_open()) { error = "Could not open system file '" + filename + "' for reading.\n"; return false; }
	std::string line;
	while (std::getline(ifs, line)) {
		trim(line);
		std::istringstream iss(line);
		std::string name, val;
		if (iss >> name >> val) { environmentSubstitute(val); map_[name] = val; }
	}
	return true;
}

double InputMap::get(const std::string& key, double defaultVal, bool* found) const
{
	auto it = map_.find(key);
	if (found) *found = it != map_.end();
	return it == map_.end() ? defaultVal : atof(it->second.c_str());
}

Vec3 InputMap::getVector(const std::string& key, const Vec3& defaultVal, bool* found) const
{
	auto it = map_.find(key);
	if (found) *found = it != map_.end();
	if (it == map_.end()) return defaultVal;
	Vec3 r{ { 0, 0, 0 } };
	std::istringstream iss(it->second);
	for (int k = 0; k < 3; k++) { std::string tok; std::getline(iss, tok, ','); r[k] = atof(tok.c_str()); }
	return r;
}

std::string InputMap::getString(const std::string& key, const std::string& defaultVal, bool* found) const
{
	auto it = map_.find(key);
	if (found) *found = it != map_.end();
	return it == map_.end() ? defaultVal : it->second;
}

} // namespace sarwhost
