// xml_lite.h -- minimal XML DOM sufficient for the GenCase case files (Case.xml).
// The reference reads them with tinyxml through JXml (src/source/JXml.cpp); only element names,
// attributes and nesting matter to the solver, which is all this parser keeps.
#pragma once
#include <string>
#include <vector>
#include <map>
#include <memory>

namespace b2s {

struct XmlNode {
  std::string name;
  std::vector<std::pair<std::string,std::string>> attrs;
  std::vector<std::unique_ptr<XmlNode>> children;

  bool has(const std::string& a)const{ for(auto& kv:attrs)if(kv.first==a)return true; return false; }
  std::string attr(const std::string& a,const std::string& def="")const{ for(auto& kv:attrs)if(kv.first==a)return kv.second; return def; }
  const XmlNode* child(const std::string& n)const{ for(auto& c:children)if(c->name==n)return c.get(); return nullptr; }
  std::vector<const XmlNode*> childs(const std::string& n)const{ std::vector<const XmlNode*> v; for(auto& c:children)if(c->name==n)v.push_back(c.get()); return v; }
  // path "a.b.c" relative to this node
  const XmlNode* path(const std::string& p)const;
};

// Parses text; returns the root element or throws std::runtime_error with a position message.
std::unique_ptr<XmlNode> xml_parse(const std::string& text);
std::unique_ptr<XmlNode> xml_load(const std::string& file);

} // namespace b2s
