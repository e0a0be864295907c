// Exercises include/claxon_b200.hpp (the C++ mirror of claxon's FlacReader / FrameReader / Block API) against
// a FLAC file: prints "<channels> <bits> <frames> <samples> <sum of samples> <xor of samples>" for the
// frame-by-frame route and for the batched route; tests/test_cpp_facade.py compares them with the goldens.
#include <cstdio>
#include <cstdlib>

#include "claxon_b200.hpp"

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    try {
        claxon::Context ctx(0);
        for (int batched = 0; batched < 2; batched++) {
            claxon::FlacReader reader = claxon::FlacReader::open(ctx, argv[1]);
            const clx_streaminfo& si = reader.streaminfo();
            unsigned long long frames = 0, samples = 0;
            long long sum = 0;
            int x = 0;
            std::vector<int32_t> buffer;
            auto eat = [&](const claxon::Block& b) {
                frames++;
                samples += b.len();
                for (uint32_t c = 0; c < b.channels(); c++)
                    for (uint32_t i = 0; i < b.duration(); i++) { sum += b.sample(c, i); x ^= b.channel(c)[i]; }
            };
            if (!batched) {
                while (auto blk = reader.blocks().read_next_or_eof(std::move(buffer))) {
                    eat(*blk);
                    buffer = std::move(*blk).into_buffer();  // recycle, as claxon users do
                }
            } else {
                for (;;) {
                    std::vector<claxon::Block> blocks = reader.blocks().read_batch(64);
                    if (blocks.empty()) break;
                    for (const auto& b : blocks) eat(b);
                }
            }
            std::printf("%u %u %llu %llu %lld %d\n", si.channels, si.bits_per_sample, frames, samples, sum, x);
        }
        // Error semantics: garbage is a FormatError that compares equal to claxon's own
        const uint8_t junk[4] = {1, 2, 3, 4};
        claxon::FrameReader fr(ctx, junk, sizeof junk);
        try {
            fr.read_next_or_eof({});
            return 3;
        } catch (const claxon::Error& e) {
            if (!(e == claxon::Error(CLX_ERR_SYNC_MISSING)) || e.kind() != claxon::Error::FormatError) return 4;
        }
    } catch (const std::exception& e) {
        std::fprintf(stderr, "exception: %s\n", e.what());
        return 1;
    }
    return 0;
}
