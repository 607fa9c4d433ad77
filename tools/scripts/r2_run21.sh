g++ -O3 -std=c++17 -I include -x c++ -c transfer-learning-library-for-object-detection_b200/csrc/host_sampling.cu -o /tmp/hs.o
g++ -O3 -I include tools/microbench/host_sampler_time.cpp /tmp/hs.o -o /tmp/hs_time && /tmp/hs_time
lscpu | grep -E "Model name|MHz|^CPU\(s\)" | head -5
