// binds a GlslScene to the globals of a translated shader that includes Scene.glsl (used inside the shader's namespace)
struct SceneBinding {
	std::vector<Material> materials; // re-packed: the buffer's std430 stride is 64 bytes, the C++ struct is 60
	void bind(const GlslScene *sc) {
		uVertices = (Vertex *)sc->vertices, uVertexIndices = (uint *)sc->vertex_indices, uTexcoords = (vec2 *)sc->texcoords;
		uTexcoordIndices = (uint *)sc->texcoord_indices, uMaterialIDs = (uint *)sc->material_ids, uTransforms = (mat3x4 *)sc->transforms;
		static_assert(sizeof(Material) == 60 && sizeof(Vertex) == 12 && sizeof(mat3x4) == 48, "layout");
		materials.resize(sc->material_count);
		for (uint32_t i = 0; i < sc->material_count; ++i)
			std::memcpy(&materials[i], (const uint8_t *)sc->materials + 64 * (size_t)i, 60);
		uMaterials = materials.data();
		for (uint32_t t = 0; t < sc->texture_count && t < kTextureNum; ++t) {
			uTextures[t].texels = (const uint8_t *)sc->textures[t].texels_rgba8, uTextures[t].width = (int)sc->textures[t].width;
			uTextures[t].height = (int)sc->textures[t].height, uTextures[t].srgb = true, uTextures[t].repeat = true;
		}
	}
};
