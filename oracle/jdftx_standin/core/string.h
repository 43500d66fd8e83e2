// Stand-in for JDFTx core/string.h — see core/vector3.h in this directory for the why.
#ifndef STANDIN_CORE_STRING_H
#define STANDIN_CORE_STRING_H
#include <string>
#include <sstream>
#include <cstring>
using std::string;
using std::istringstream;
using std::ostringstream;

//! strip leading/trailing whitespace in place (InputMap.cpp:33)
inline void trim(string& s)
{	const char* ws = " \t\r\n\v\f";
	size_t b = s.find_first_not_of(ws);
	if(b == string::npos) { s.clear(); return; }
	size_t e = s.find_last_not_of(ws);
	s = s.substr(b, e - b + 1);
}
#endif
