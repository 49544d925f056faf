cd "$(dirname "$0")"
[ -x ./tma_bw ] || nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tma_bw tma_bw.cu -lcuda
# box bi br, tile twf th, S, ctas/SM, xoff
./tma_bw 192 40  96 24 3 2 0
./tma_bw 192 40  96 24 3 2 -12
./tma_bw 192 40  96 24 6 1 -12
./tma_bw 192 40  96 24 2 3 -12
./tma_bw 128 24  96 24 3 2 0
./tma_bw 96 24   96 24 4 2 0
./tma_bw 96 24   96 24 8 4 0
./tma_bw 256 36  192 24 2 2 -12
./tma_bw 256 36  192 24 1 4 -12
./tma_bw 256 36  192 24 4 1 -12
./tma_bw 192 24  192 24 4 2 0
./tma_bw 192 24  192 24 8 2 0
./tma_bw 256 32  256 32 3 2 0
./tma_bw 256 32  256 32 6 1 0
./tma_bw 32 40   96 24 3 2 0
