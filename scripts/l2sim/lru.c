// LRU simulation of row fetches of a CSR gather. usage: lru rowptr.bin col.bin order.bin n cap
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
int main(int argc,char**argv){
  int n=atoi(argv[4]); 
  FILE*f=fopen(argv[1],"rb"); int64_t*rp=malloc(8*(n+1)); if(fread(rp,8,n+1,f)!=(size_t)n+1)return 1; fclose(f);
  int64_t nnz=rp[n]; int32_t*col=malloc(4*nnz); f=fopen(argv[2],"rb"); if(fread(col,4,nnz,f)!=(size_t)nnz)return 1; fclose(f);
  int32_t*ord=malloc(4*n); f=fopen(argv[3],"rb"); if(fread(ord,4,n,f)!=(size_t)n)return 1; fclose(f);
  for(int a=5;a<argc;a++){
    int cap=atoi(argv[a]);
    int32_t*prev=malloc(4*(n+1)),*next=malloc(4*(n+1)); char*in=calloc(n,1);
    int head=n; prev[head]=next[head]=head; int cnt=0; int64_t miss=0,acc=0;
    for(int k=0;k<n;k++){int r=ord[k];
      for(int64_t e=rp[r];e<rp[r+1];e++){int c=col[e];acc++;
        if(in[c]){ // unlink
          next[prev[c]]=next[c]; prev[next[c]]=prev[c];
        } else { miss++; in[c]=1; cnt++;
          if(cnt>cap){int v=prev[head]; next[prev[v]]=head; prev[head]=prev[v]; in[v]=0; cnt--;}
        }
        // push front
        next[c]=next[head]; prev[c]=head; prev[next[head]]=c; next[head]=c;
      }}
    printf("cap %d: accesses %ld misses %ld hit %.3f\n",cap,(long)acc,(long)miss,1.0-(double)miss/acc);
    free(prev);free(next);free(in);
  }
}
