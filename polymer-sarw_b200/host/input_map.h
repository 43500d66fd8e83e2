// input_map.h — the reference's config dictionary semantics (InputMap.h:28-35, InputMap.cpp:27-86), rebuilt without JDFTx.
//   * one "name value" pair per line: the line is trimmed, the first two whitespace-separated tokens are taken, the
//     rest of the line is ignored (inline comments), later keys override earlier ones        InputMap.cpp:31-43
//   * ${VAR} / $VAR in values are replaced from the environment                              InputMap.cpp:38-40
//   * get(): atof of the value; getVector(): three comma-separated atof()s; getString()      InputMap.cpp:47-86
//   * a missing key without default is fatal: message + exit(1)                              InputMap.cpp:49-52,60-63,79-82
//   * unknown keys are stored and silently ignored                                           InputMap.cpp:41
#pragma once
#include <map>
#include <string>

namespace sarwhost {

struct Vec3 { double v[3]; double& operator[](int k) { return v[k]; } const double& operator[](int k) const { return v[k]; } };

class InputMap {
public:
	// returns false (and sets error) when the file cannot be opened (the reference dies: InputMap.cpp:29-30)
	bool load(const std::string& filename, std::string& error);
	bool has(const std::string& key) const { return map_.count(key) != 0; }
	// `found` reports whether the key was present (the caller dies for required keys, as the reference does)
	double get(const std::string& key, double defaultVal, bool* found = nullptr) const;
	Vec3 getVector(const std::string& key, const Vec3& defaultVal, bool* found = nullptr) const;
	std::string getString(const std::string& key, const std::string& defaultVal, bool* found = nullptr) const;
	const std::map<std::string, std::string>& entries() const { return map_; }
private:
	std::map<std::string, std::string> map_;
};

void trim(std::string& s);
void environmentSubstitute(std::string& s);

} // namespace sarwhost
