#include "../../sgufp_solver_b200/csrc/model.cpp"
using namespace sgufp;
extern "C" int plan_graph(int n,int m,const int32_t*tail,const int32_t*head,const int32_t*rew0,const int32_t*vbar,int nvbar,
                          const int16_t*path,int L,int32_t*sv,int32_t*ev,int32_t*r,int32_t*nc_out,int32_t*arc_chain){
    Model M; std::string err;
    if(int rc=M.build(n,m,tail,head,rew0,vbar,nvbar,err)) return rc;
    Plan P; if(int rc=build_plan(M,path,L,P,err)) return rc;
    const PlanHeader*h=reinterpret_cast<const PlanHeader*>(P.words.data());
    const int32_t*w=P.words.data();
    for(int c=0;c<h->nopen;c++){int e=w[h->o_ch_ends+c]; sv[c]=(e&0xffff)-1; ev[c]=(e>>16)-1; r[c]=w[h->o_ch_r+c];}
    for(int a=0;a<m;a++){int cp=w[h->o_arc_cp+a]; arc_chain[a]=(cp>>10)<h->nopen?(cp>>10):-1;}
    *nc_out=h->nc; return h->nopen;
}
