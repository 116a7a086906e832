// obj_loader.cpp -- CLOBJloader: the OBJ/MTL dialect the reference accepts (CLOBJloader.cpp:10-176),
// parsed from an in-memory copy of the file with a small cursor instead of fscanf/fgets/strtok.
//
// Behaviour kept from the reference, because it defines the triangle array and therefore hit IDs:
//   * the stream is a sequence of whitespace-separated words; only `v`, `vt`, `vn`, `usemtl`, `f`
//     (OBJ) and `newmtl`, `Kd`, `Ks`, `Ke`, `Ns`, `Ni` (MTL) are acted on, every other word --
//     comments included -- is skipped one word at a time (CLOBJloader.cpp:39-46, 140-147);
//   * numbers are read the way scanf("%f") reads them (strtof after skipping white space; a field
//     that does not parse leaves the value unchanged and the cursor in place);
//   * an `f` record is the rest of its line, at most 127 characters, split on SPACES only; words of
//     length <= 1 are ignored; each word must be `v/vt/vn` (CLOBJloader.cpp:79-100);
//   * an n-vertex face yields triangles (i, i+1, i+2) for i = 0..n-3 and then (n-2, n-1, 0): a
//     triangular face therefore appears TWICE, the second copy with rotated vertices
//     (CLOBJloader.cpp:102-126);
//   * `usemtl` with an unknown name keeps the previous material index; the index starts at
//     0xFFFFFFFF (CLOBJloader.cpp:37, 66-77);
//   * the material file is the scene path with its last four characters replaced by ".mtl".
// Deliberate differences: malformed input that makes the reference read out of bounds (missing
// v/vt/vn fields, indices outside the arrays, a material statement before any `newmtl`, paths longer
// than its 80-byte buffer are accepted) raises CLException here instead.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "glaze3d.h"

namespace Glaze3D
{
    namespace
    {
        struct Cursor
        {
            const char* p;
            const char* end;
            static bool space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\v' || c == '\f'; }
            void skipSpace() { while (p < end && space(*p)) ++p; }
            // scanf("%s") with a width limit: next run of non-space characters
            bool word(std::string& out, size_t limit = 127)
            {
                skipSpace();
                if (p >= end) return false;
                const char* s = p;
                while (p < end && !space(*p) && (size_t)(p - s) < limit) ++p;
                out.assign(s, p);
                return true;
            }
            // scanf("%f"): on failure nothing is consumed beyond the leading white space. strtof does the work in
            // the general case; plain decimals (what OBJ writers emit) take a fast path that returns the SAME float:
            // up to 15 significant digits and a power of ten up to 10^22 convert exactly to a correctly rounded
            // double (Clinger's fast path), and rounding that double to float equals the correctly rounded float
            // unless the double sits exactly on a float rounding boundary -- the one case handed back to strtof.
            bool number(float& out)
            {
                skipSpace();
                if (p >= end) return false;
                const char* q = p;
                bool neg = false;
                if (*q == '+' || *q == '-') { neg = *q == '-'; ++q; }
                uint64_t mant = 0;
                int digits = 0, exp10 = 0;
                bool any = false, simple = true;
                if (q + 1 < end && q[0] == '0' && (q[1] == 'x' || q[1] == 'X')) simple = false;    // hex float: strtof's business
                while (q < end && *q >= '0' && *q <= '9')
                {
                    any = true;
                    if (mant || *q != '0') { if (digits < 15) { mant = mant * 10 + (uint64_t)(*q - '0'); ++digits; } else simple = false; }
                    ++q;
                }
                if (q < end && *q == '.')
                {
                    ++q;
                    while (q < end && *q >= '0' && *q <= '9')
                    {
                        any = true;
                        if (mant || *q != '0') { if (digits < 15) { mant = mant * 10 + (uint64_t)(*q - '0'); ++digits; --exp10; } else simple = false; }
                        else --exp10;
                        ++q;
                    }
                }
                if (any && q < end && (*q == 'e' || *q == 'E'))
                {
                    const char* e = q + 1;
                    bool eneg = false;
                    if (e < end && (*e == '+' || *e == '-')) { eneg = *e == '-'; ++e; }
                    if (e < end && *e >= '0' && *e <= '9')
                    {
                        int ev = 0;
                        while (e < end && *e >= '0' && *e <= '9') { if (ev < 10000) ev = ev * 10 + (*e - '0'); ++e; }
                        exp10 += eneg ? -ev : ev;
                        q = e;
                    }
                }
                if (any && simple && exp10 >= -22 && exp10 <= 22)
                {
                    static const double pow10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                                      1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };
                    double d = (double)mant;
                    d = exp10 < 0 ? d / pow10[-exp10] : d * pow10[exp10];
                    uint64_t bits;
                    std::memcpy(&bits, &d, 8);
                    const double mag = d;
                    if ((bits & 0x1FFFFFFFull) != 0x10000000ull && (mant == 0 || (mag > 1e-30 && mag < 1e30)))
                    {
                        float f = (float)d;
                        out = neg ? -f : f;
                        p = q;
                        return true;
                    }
                }
                char* stop = nullptr;
                float v = std::strtof(p, &stop);
                if (stop == p) return false;
                out = v;
                p = stop;
                return true;
            }
            void numbers(float* dst, int n) { for (int i = 0; i < n; ++i) if (!number(dst[i])) return; }
            // fgets(buf, 128): up to 127 characters, through the newline
            std::string restOfLine()
            {
                const char* s = p;
                while (p < end && (p - s) < 127) { if (*p++ == '\n') break; }
                return std::string(s, p);
            }
        };

        std::string slurp(const std::string& path, const char* what)
        {
            FILE* f = std::fopen(path.c_str(), "rb");
            if (!f) throw CLException(std::string("Failed to open ") + what + " file '" + path + "'", B2RT_INVALID_VALUE);
            std::string data;
            char buf[1 << 16];
            size_t n;
            while ((n = std::fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, n);
            std::fclose(f);
            data.push_back('\0');     // strtof needs a terminator
            return data;
        }

        void loadMaterials(CLBVHScene& scene, const std::string& path)
        {
            std::string data = slurp(path, "material");
            Cursor c{ data.data(), data.data() + data.size() - 1 };
            std::string w;
            auto current = [&]() -> CLMaterial& {
                if (scene.m_Materials.empty()) throw CLException("Material statement before any newmtl in '" + path + "'", B2RT_INVALID_VALUE);
                return scene.m_Materials.back();
            };
            while (c.word(w))
            {
                if (w == "newmtl")
                {
                    std::string name;
                    c.word(name, 79);
                    scene.m_MaterialNames.push_back(name);
                    scene.m_Materials.push_back(CLMaterial());
                }
                else if (w == "Kd") c.numbers(&current().diffuse.x, 3);
                else if (w == "Ks") c.numbers(&current().specular.x, 3);
                else if (w == "Ke") c.numbers(&current().emission.x, 3);
                else if (w == "Ns") c.numbers(&current().roughness, 1);
                else if (w == "Ni") c.numbers(&current().ior, 1);
            }
        }
    }

    namespace
    {
        // sscanf(token, "%d/%d/%d", &a, &b, &n) (CLOBJloader.cpp:96) without the library call: each %d skips white space, takes
        // an optional sign and at least one digit; a conversion or a '/' that does not match ends the scan and leaves the
        // remaining outputs untouched (0). Values wrap like the reference's int -> unsigned int assignment, out-of-range
        // digit strings like glibc's.
        bool scanInt(const char*& p, const char* end, unsigned int& out)
        {
            const char* q = p;
            while (q < end && (*q == ' ' || (*q >= '\t' && *q <= '\r'))) ++q;
            bool neg = false;
            if (q < end && (*q == '+' || *q == '-')) { neg = *q == '-'; ++q; }
            if (q >= end || *q < '0' || *q > '9') return false;
            // glibc reads %d through a 64-bit strtol (saturating at LONG_MAX / LONG_MIN) and stores the low 32 bits
            const unsigned long long limit = neg ? (1ull << 63) : (1ull << 63) - 1;
            unsigned long long v = 0;
            bool sat = false;
            while (q < end && *q >= '0' && *q <= '9')
            {
                const unsigned d = (unsigned)(*q - '0');
                if (!sat && (v > (limit - d) / 10)) sat = true;
                if (!sat) v = v * 10 + d;
                ++q;
            }
            if (sat) v = limit;
            out = neg ? (unsigned int)(0ull - v) : (unsigned int)v;
            p = q;
            return true;
        }
        void scanTriplet(const char* p, const char* end, unsigned int& a, unsigned int& b, unsigned int& n)
        {
            if (!scanInt(p, end, a)) return;
            if (p >= end || *p != '/') return;
            ++p;
            if (!scanInt(p, end, b)) return;
            if (p >= end || *p != '/') return;
            ++p;
            scanInt(p, end, n);
        }
    }

#ifdef G3D_TEST_HOOKS
    // TEST BUILD ONLY (the test harness compiles this file with -DG3D_TEST_HOOKS into a library of its own): the scanners
    // behind the loader, reachable one token at a time. The product library does not contain them.
    // Test hook: one face-vertex token through the triplet scanner.
    void ScanTripletForTest(const char* token, unsigned int out[3])
    {
        out[0] = out[1] = out[2] = 0;
        scanTriplet(token, token + std::strlen(token), out[0], out[1], out[2]);
    }

    // Test hook: the numbers of `text` as the loader reads them (scanf("%f") semantics), up to maxCount.
    int ParseNumbersForTest(const char* text, float* out, int maxCount)
    {
        std::string data(text ? text : "");
        data.push_back('\0');
        Cursor c{ data.data(), data.data() + data.size() - 1 };
        int n = 0;
        while (n < maxCount && c.number(out[n])) ++n;
        return n;
    }
#endif

    void CLOBJloader::LoadInto(CLBVHScene& scene, const char* filename)
    {
        std::string path(filename ? filename : "");
        // The reference copies the path into an 80-byte buffer and overflows it beyond ~75 characters (CLOBJloader.cpp:18-23);
        // longer paths are simply accepted here.
        if (path.size() < 5) throw CLException("Scene path must end in a 4-character extension", B2RT_INVALID_VALUE);
        loadMaterials(scene, path.substr(0, path.size() - 4) + ".mtl");

        std::string data = slurp(path, "scene");
        Cursor c{ data.data(), data.data() + data.size() - 1 };
        std::vector<float3> positions, normals;
        std::vector<float2> texcoords;
        {   // size the arrays once: count the lines by their first two characters (a hint only, the parse below decides)
            size_t nf = 0, nv = 0, nn = 0, nt = 0;
            const char* e = data.data() + data.size() - 1;
            for (const char* q = data.data(); q && q + 2 < e; q = static_cast<const char*>(std::memchr(q, '\n', (size_t)(e - q))), q = q ? q + 1 : nullptr)
            {
                if (q[0] == 'f' && q[1] == ' ') ++nf;
                else if (q[0] == 'v' && q[1] == ' ') ++nv;
                else if (q[0] == 'v' && q[1] == 'n') ++nn;
                else if (q[0] == 'v' && q[1] == 't') ++nt;
            }
            positions.reserve(nv); normals.reserve(nn); texcoords.reserve(nt);
            scene.m_Triangles.reserve(scene.m_Triangles.size() + 2 * nf);
        }
        unsigned int materialIndex = 0xFFFFFFFFu;
        std::string w;
        std::vector<unsigned int> iv, it, in;

        auto vertex = [&](size_t k) {
            unsigned int a = iv[k] - 1, b = it[k] - 1, n = in[k] - 1;     // 0 (missing field) wraps to 0xFFFFFFFF
            if (a >= positions.size() || b >= texcoords.size() || n >= normals.size())
                throw CLException("Face references v/vt/vn " + std::to_string(iv[k]) + "/" + std::to_string(it[k]) + "/" + std::to_string(in[k]) +
                                  " outside the arrays read so far (faces must be v/vt/vn triplets with positive indices)", B2RT_INVALID_VALUE);
            return CLVertex(positions[a], texcoords[b], normals[n]);
        };

        while (c.word(w))
        {
            if (w == "v") { float3 p; c.numbers(&p.x, 3); positions.push_back(p); }
            else if (w == "vt") { float2 t; c.numbers(&t.x, 2); texcoords.push_back(t); }
            else if (w == "vn") { float3 n; c.numbers(&n.x, 3); normals.push_back(n); }
            else if (w == "usemtl")
            {
                std::string name;
                c.word(name, 79);
                for (unsigned int i = 0; i < scene.m_MaterialNames.size(); ++i)
                    if (scene.m_MaterialNames[i] == name) { materialIndex = i; break; }
            }
            else if (w == "f")
            {
                std::string line = c.restOfLine();
                iv.clear(); it.clear(); in.clear();
                size_t pos = 0;
                while (pos < line.size())
                {
                    size_t stop = line.find(' ', pos);
                    if (stop == std::string::npos) stop = line.size();
                    if (stop - pos > 1)
                    {
                        unsigned int a = 0, b = 0, n = 0;
                        scanTriplet(line.data() + pos, line.data() + stop, a, b, n);
                        iv.push_back(a); it.push_back(b); in.push_back(n);
                    }
                    pos = stop + 1;
                }
                if (iv.size() < 2)
                {
                    if (iv.empty()) continue;
                    throw CLException("Face with a single vertex", B2RT_INVALID_VALUE);
                }
                for (size_t i = 0; i + 2 < iv.size(); ++i)
                    scene.m_Triangles.push_back(CLTriangle(vertex(i), vertex(i + 1), vertex(i + 2), materialIndex));
                scene.m_Triangles.push_back(CLTriangle(vertex(iv.size() - 2), vertex(iv.size() - 1), vertex(0), materialIndex));
            }
        }
    }

    void CLOBJloader::Load(const char* filename, unsigned int maxPrimitivesInNode)
    {
        if (!eng || !eng->render || !eng->render->m_Scene)
            throw CLException("CLOBJloader::Load needs eng->render->m_Scene (the reference writes into the global engine)", B2RT_INVALID_CONTEXT);
        eng->render->m_Scene->m_MaxPrimitivesInNode = maxPrimitivesInNode;
        LoadInto(*eng->render->m_Scene, filename);
    }
}

#ifdef G3D_TEST_HOOKS
extern "C" void g3d_scan_triplet(const char* token, unsigned int* out) { Glaze3D::ScanTripletForTest(token, out); }
extern "C" int g3d_parse_numbers(const char* text, float* out, int maxCount) { return Glaze3D::ParseNumbersForTest(text, out, maxCount); }
#endif
