// scene_cache.cpp -- binary cache of the post-loader, post-build scene arrays.
//
// The reference parses the OBJ text with fscanf and rebuilds the SAH BVH on every start
// (CLEngineBase.cpp:172-179 -> CLOBJloader.cpp:16-129, CLBVHnode.cpp:185-207): ~1.3 s per 24 MB of OBJ text plus
// ~1.4 s per 655 k triangles, single-threaded. The cache stores exactly what those two steps produce --
// m_Triangles in their post-build order (which defines hit IDs, CLBVHnode.cpp:197), the flattened
// CLLinearBVHNode array, the materials and their names -- so a cache hit gives byte-identical arrays and
// therefore identical IDs. A cache file is tied to its source by the size and modification time of the .obj and
// .mtl files, the maxPrimitivesInNode it was built with, the record sizes and a checksum of the payload; anything
// that does not match is treated as a miss and the cache is rewritten after a normal load + build.
#include <sys/stat.h>
#include <unistd.h>
#include <cstdio>
#include <cstring>
#include "glaze3d.h"

namespace Glaze3D
{
    namespace
    {
        struct CacheHeader
        {
            char magic[8];                 // "B2RTSCN" + version byte
            uint32_t maxPrims, recordSizes; // recordSizes = 256 | 48 << 12 | 64 << 20
            uint64_t nTris, nNodes, nMats, nameBytes;
            uint64_t objSize, mtlSize;
            int64_t objMtimeNs, mtlMtimeNs;
            uint64_t checksum;
        };
        const char MAGIC[8] = { 'B', '2', 'R', 'T', 'S', 'C', 'N', 1 };
        const uint32_t RECORD_SIZES = 256u | (48u << 12) | (64u << 20);

        bool stamp(const std::string& path, uint64_t& size, int64_t& mtimeNs)
        {
            struct stat st;
            if (::stat(path.c_str(), &st) != 0) return false;
            size = (uint64_t)st.st_size;
            mtimeNs = (int64_t)st.st_mtim.tv_sec * 1000000000ll + st.st_mtim.tv_nsec;
            return true;
        }
        std::string mtlPath(const std::string& obj) { return obj.size() >= 4 ? obj.substr(0, obj.size() - 4) + ".mtl" : obj; }

        // Four interleaved multiply-rotate lanes over 64-bit words: runs at memory speed, catches truncation and bit rot.
        uint64_t mix(uint64_t h, const void* data, size_t bytes)
        {
            const unsigned char* p = static_cast<const unsigned char*>(data);
            uint64_t lane[4] = { h, h ^ 0x9E3779B97F4A7C15ull, h + 0xC2B2AE3D27D4EB4Full, ~h };
            size_t words = bytes / 8, i = 0;
            for (; i + 4 <= words; i += 4)
                for (int k = 0; k < 4; ++k)
                {
                    uint64_t w;
                    std::memcpy(&w, p + 8 * (i + k), 8);
                    lane[k] = ((lane[k] ^ w) * 0x100000001B3ull);
                    lane[k] = (lane[k] << 29) | (lane[k] >> 35);
                }
            uint64_t out = lane[0] ^ (lane[1] * 3) ^ (lane[2] * 5) ^ (lane[3] * 7);
            for (size_t b = 8 * i; b < bytes; ++b) out = (out ^ p[b]) * 0x100000001B3ull;
            return out;
        }

        struct File
        {
            FILE* f = nullptr;
            ~File() { if (f) std::fclose(f); }
        };
    }

    std::string CLBVHScene::CachePathFor(const char* objPath, unsigned int maxPrimitivesInNode)
    {
        return std::string(objPath ? objPath : "") + ".p" + std::to_string(maxPrimitivesInNode) + ".b2rtscn";
    }

    void CLBVHScene::SaveCache(const char* cachePath, const char* objPath) const
    {
        if (m_Nodes.empty() || m_Triangles.empty()) throw CLException("SaveCache needs a built scene", B2RT_INVALID_VALUE);
        CacheHeader h;
        std::memset(&h, 0, sizeof(h));
        std::memcpy(h.magic, MAGIC, 8);
        h.maxPrims = m_MaxPrimitivesInNode;
        h.recordSizes = RECORD_SIZES;
        h.nTris = m_Triangles.size(); h.nNodes = m_Nodes.size(); h.nMats = m_Materials.size();
        std::string names;
        for (const std::string& n : m_MaterialNames) { names += n; names.push_back('\0'); }
        h.nameBytes = names.size();
        std::string obj(objPath ? objPath : "");
        if (!stamp(obj, h.objSize, h.objMtimeNs) || !stamp(mtlPath(obj), h.mtlSize, h.mtlMtimeNs))
            throw CLException("SaveCache: cannot stat '" + obj + "' and its .mtl", B2RT_INVALID_VALUE);
        uint64_t c = mix(0x5CE9Eull, m_Triangles.data(), m_Triangles.size() * sizeof(CLTriangle));
        c = mix(c, m_Nodes.data(), m_Nodes.size() * sizeof(CLLinearBVHNode));
        c = mix(c, m_Materials.data(), m_Materials.size() * sizeof(CLMaterial));
        h.checksum = mix(c, names.data(), names.size());

        // write beside the target and rename: readers never see a half-written cache
        std::string tmp = std::string(cachePath) + ".tmp." + std::to_string((long long)::getpid());
        {
            File out;
            out.f = std::fopen(tmp.c_str(), "wb");
            if (!out.f) throw CLException("SaveCache: cannot create '" + tmp + "'", B2RT_INVALID_VALUE);
            bool ok = std::fwrite(&h, sizeof(h), 1, out.f) == 1;
            ok = ok && std::fwrite(m_Triangles.data(), sizeof(CLTriangle), m_Triangles.size(), out.f) == m_Triangles.size();
            ok = ok && std::fwrite(m_Nodes.data(), sizeof(CLLinearBVHNode), m_Nodes.size(), out.f) == m_Nodes.size();
            ok = ok && (m_Materials.empty() || std::fwrite(m_Materials.data(), sizeof(CLMaterial), m_Materials.size(), out.f) == m_Materials.size());
            ok = ok && (names.empty() || std::fwrite(names.data(), 1, names.size(), out.f) == names.size());
            ok = ok && std::fflush(out.f) == 0;
            if (!ok) { std::fclose(out.f); out.f = nullptr; std::remove(tmp.c_str()); throw CLException("SaveCache: short write to '" + tmp + "'", B2RT_OUT_OF_RESOURCES); }
        }
        if (std::rename(tmp.c_str(), cachePath) != 0) { std::remove(tmp.c_str()); throw CLException(std::string("SaveCache: cannot rename to '") + cachePath + "'", B2RT_INVALID_VALUE); }
    }

    bool CLBVHScene::LoadCache(const char* cachePath, const char* objPath, unsigned int maxPrimitivesInNode)
    {
        File in;
        in.f = std::fopen(cachePath, "rb");
        if (!in.f) return false;
        CacheHeader h;
        if (std::fread(&h, sizeof(h), 1, in.f) != 1) return false;
        if (std::memcmp(h.magic, MAGIC, 8) != 0 || h.recordSizes != RECORD_SIZES || h.maxPrims != maxPrimitivesInNode) return false;
        std::string obj(objPath ? objPath : "");
        uint64_t objSize = 0, mtlSize = 0;
        int64_t objM = 0, mtlM = 0;
        if (!stamp(obj, objSize, objM) || !stamp(mtlPath(obj), mtlSize, mtlM)) return false;
        if (objSize != h.objSize || objM != h.objMtimeNs || mtlSize != h.mtlSize || mtlM != h.mtlMtimeNs) return false;   // source changed
        if (h.nTris == 0 || h.nNodes == 0 || h.nTris >= 0xfffffffeull || h.nNodes >= 0xfffffffeull || h.nMats > (1u << 24) || h.nameBytes > (1u << 28)) return false;
        struct stat st;
        if (::fstat(::fileno(in.f), &st) != 0) return false;
        const uint64_t want = sizeof(h) + h.nTris * sizeof(CLTriangle) + h.nNodes * sizeof(CLLinearBVHNode) + h.nMats * sizeof(CLMaterial) + h.nameBytes;
        if ((uint64_t)st.st_size != want) return false;                                                                  // truncated or padded

        std::vector<CLTriangle> tris(h.nTris);
        std::vector<CLLinearBVHNode> nodes(h.nNodes);
        std::vector<CLMaterial> mats(h.nMats);
        std::string names(h.nameBytes, '\0');
        if (std::fread(tris.data(), sizeof(CLTriangle), tris.size(), in.f) != tris.size()) return false;
        if (std::fread(nodes.data(), sizeof(CLLinearBVHNode), nodes.size(), in.f) != nodes.size()) return false;
        if (!mats.empty() && std::fread(mats.data(), sizeof(CLMaterial), mats.size(), in.f) != mats.size()) return false;
        if (!names.empty() && std::fread(&names[0], 1, names.size(), in.f) != names.size()) return false;
        uint64_t c = mix(0x5CE9Eull, tris.data(), tris.size() * sizeof(CLTriangle));
        c = mix(c, nodes.data(), nodes.size() * sizeof(CLLinearBVHNode));
        c = mix(c, mats.data(), mats.size() * sizeof(CLMaterial));
        if (mix(c, names.data(), names.size()) != h.checksum) return false;

        m_Triangles.swap(tris);
        m_Nodes.swap(nodes);
        m_Materials.swap(mats);
        m_MaterialNames.clear();
        for (size_t pos = 0; pos < names.size();)
        {
            size_t end = names.find('\0', pos);
            if (end == std::string::npos) end = names.size();
            m_MaterialNames.push_back(names.substr(pos, end - pos));
            pos = end + 1;
        }
        m_MaxPrimitivesInNode = maxPrimitivesInNode;
        return true;
    }

    bool CLOBJloader::LoadCached(CLBVHScene& scene, const char* filename, unsigned int maxPrimitivesInNode, const char* cachePath)
    {
        const std::string cache = cachePath && *cachePath ? std::string(cachePath) : CLBVHScene::CachePathFor(filename, maxPrimitivesInNode);
        if (scene.LoadCache(cache.c_str(), filename, maxPrimitivesInNode)) return true;
        scene.m_Triangles.clear(); scene.m_Materials.clear(); scene.m_MaterialNames.clear();
        scene.m_MaxPrimitivesInNode = maxPrimitivesInNode;
        LoadInto(scene, filename);
        scene.BuildOnly(maxPrimitivesInNode);
        try { scene.SaveCache(cache.c_str(), filename); }
        catch (const CLException&) { /* a read-only directory must not break scene loading */ }
        return false;
    }
}
