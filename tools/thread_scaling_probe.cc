#include <thread>
#include <vector>
#include <chrono>
#include <cstdio>
#include <atomic>
int main(){
  for (int nt : {1,2,4,8}) {
    std::atomic<long> sink{0};
    auto t0=std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    for(int t=0;t<nt;++t) th.emplace_back([&]{ volatile long x=0; for(long i=0;i<400000000L/nt;++i) x+=i; sink+=x;});
    for(auto&t:th) t.join();
    printf("%d threads: %.1f ms\n", nt, std::chrono::duration<double,std::milli>(std::chrono::steady_clock::now()-t0).count());
  }
}
