// xml_lite.cpp -- see xml_lite.h.
#include "xml_lite.h"
#include <stdexcept>
#include <fstream>
#include <sstream>
#include <cctype>

namespace b2s {

const XmlNode* XmlNode::path(const std::string& p)const{
  const XmlNode* n=this;
  size_t i=0;
  while(n && i<p.size()){
    size_t j=p.find('.',i);
    if(j==std::string::npos)j=p.size();
    n=n->child(p.substr(i,j-i));
    i=j+1;
  }
  return n;
}

namespace {
struct Parser {
  const std::string& s; size_t i;
  explicit Parser(const std::string& t):s(t),i(0){}
  [[noreturn]] void fail(const char* m){ throw std::runtime_error(std::string("xml: ")+m+" at offset "+std::to_string(i)); }
  void ws(){ while(i<s.size() && isspace((unsigned char)s[i]))i++; }
  bool starts(const char* t){ size_t n=0; while(t[n])n++; return s.compare(i,n,t)==0; }
  void skip_misc(){
    for(;;){
      ws();
      if(starts("<?")){ size_t e=s.find("?>",i); if(e==std::string::npos)fail("unterminated <?"); i=e+2; }
      else if(starts("<!--")){ size_t e=s.find("-->",i); if(e==std::string::npos)fail("unterminated comment"); i=e+3; }
      else if(starts("<!")){ size_t e=s.find('>',i); if(e==std::string::npos)fail("unterminated <!"); i=e+1; }
      else break;
    }
  }
  std::string name(){ size_t b=i; while(i<s.size() && (isalnum((unsigned char)s[i])||s[i]=='_'||s[i]=='-'||s[i]==':'||s[i]=='.'))i++; if(b==i)fail("name expected"); return s.substr(b,i-b); }
  static std::string unescape(const std::string& v){
    std::string o; o.reserve(v.size());
    for(size_t k=0;k<v.size();k++){
      if(v[k]=='&'){
        if(v.compare(k,4,"&lt;")==0){ o+='<'; k+=3; }
        else if(v.compare(k,4,"&gt;")==0){ o+='>'; k+=3; }
        else if(v.compare(k,5,"&amp;")==0){ o+='&'; k+=4; }
        else if(v.compare(k,6,"&quot;")==0){ o+='"'; k+=5; }
        else if(v.compare(k,6,"&apos;")==0){ o+='\''; k+=5; }
        else o+=v[k];
      }
      else o+=v[k];
    }
    return o;
  }
  std::unique_ptr<XmlNode> element(){
    if(s[i]!='<')fail("'<' expected");
    i++;
    std::unique_ptr<XmlNode> n(new XmlNode());
    n->name=name();
    for(;;){
      ws();
      if(i>=s.size())fail("unexpected end");
      if(s[i]=='/'){ if(i+1>=s.size()||s[i+1]!='>')fail("'/>' expected"); i+=2; return n; }
      if(s[i]=='>'){ i++; break; }
      std::string a=name(); ws();
      if(i>=s.size()||s[i]!='=')fail("'=' expected");
      i++; ws();
      if(i>=s.size()||(s[i]!='"'&&s[i]!='\''))fail("quote expected");
      const char q=s[i++]; size_t e=s.find(q,i); if(e==std::string::npos)fail("unterminated attribute");
      n->attrs.emplace_back(a,unescape(s.substr(i,e-i))); i=e+1;
    }
    for(;;){
      // text is ignored
      size_t lt=s.find('<',i); if(lt==std::string::npos)fail("unterminated element");
      i=lt;
      if(starts("</")){ i+=2; std::string c=name(); if(c!=n->name)fail("mismatched closing tag"); ws(); if(i>=s.size()||s[i]!='>')fail("'>' expected"); i++; return n; }
      if(starts("<!--")){ size_t e=s.find("-->",i); if(e==std::string::npos)fail("unterminated comment"); i=e+3; continue; }
      if(starts("<?")){ size_t e=s.find("?>",i); if(e==std::string::npos)fail("unterminated <?"); i=e+2; continue; }
      if(starts("<![CDATA[")){ size_t e=s.find("]]>",i); if(e==std::string::npos)fail("unterminated CDATA"); i=e+3; continue; }
      n->children.push_back(element());
    }
  }
};
}

std::unique_ptr<XmlNode> xml_parse(const std::string& text){
  Parser p(text);
  p.skip_misc();
  if(p.i>=text.size())throw std::runtime_error("xml: no root element");
  return p.element();
}

std::unique_ptr<XmlNode> xml_load(const std::string& file){
  std::ifstream f(file.c_str(),std::ios::binary);
  if(!f)throw std::runtime_error("xml: cannot open "+file);
  std::stringstream ss; ss<<f.rdbuf();
  return xml_parse(ss.str());
}

} // namespace b2s
